// comm.cu -- multi-GPU plumbing: one process per GPU, NCCL over NVLink.
//
// Replaces the host-staged MPI halo exchange of benchmarks/multigpu_mpi_benchmark.jl:302-360 (device -> host ->
// Isend/Irecv -> host -> device every iteration) and the MPI.Allreduce dots of demos/krylov_mpi_gpu_demo.jl:231-277.
// Here the interface values are packed on the device, exchanged with grouped ncclSend/ncclRecv on the handle's
// stream (device buffers end to end, no host copy, no host synchronisation) and received straight into the ghost
// segment of the vector (ghosts of one neighbour are contiguous because ownership is by contiguous id range).
//
// NCCL is resolved with dlopen at the first comm call, so the library has no link-time NCCL dependency and a
// process that already loaded an NCCL (e.g. through torch) shares that copy.
#include <dlfcn.h>

#include "handle.h"

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid *);
typedef int (*fn_init_rank)(ncclComm **, int, nccl_uid, int);
typedef int (*fn_destroy)(ncclComm *);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, ncclComm *, cudaStream_t);
typedef int (*fn_sendrecv)(void *, size_t, int, int, ncclComm *, cudaStream_t);
typedef int (*fn_void)(void);
typedef const char *(*fn_errstr)(int);

static struct {
    void *so = nullptr;
    fn_get_uid get_uid; fn_init_rank init_rank; fn_destroy destroy; fn_allreduce allreduce;
    fn_sendrecv send, recv; fn_void group_start, group_end; fn_errstr errstr;
} N;

#define NCCL_DOUBLE 8   // ncclFloat64
#define NCCL_SUM 0

static int nccl_load() {
    if (N.so) return JFEM_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !N.so; i++) N.so = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!N.so) { jfem_set_error("cannot load libnccl.so.2: %s", dlerror()); return JFEM_ENCCL; }
    N.get_uid = (fn_get_uid)dlsym(N.so, "ncclGetUniqueId");
    N.init_rank = (fn_init_rank)dlsym(N.so, "ncclCommInitRank");
    N.destroy = (fn_destroy)dlsym(N.so, "ncclCommDestroy");
    N.allreduce = (fn_allreduce)dlsym(N.so, "ncclAllReduce");
    N.send = (fn_sendrecv)dlsym(N.so, "ncclSend");
    N.recv = (fn_sendrecv)dlsym(N.so, "ncclRecv");
    N.group_start = (fn_void)dlsym(N.so, "ncclGroupStart");
    N.group_end = (fn_void)dlsym(N.so, "ncclGroupEnd");
    N.errstr = (fn_errstr)dlsym(N.so, "ncclGetErrorString");
    if (!N.get_uid || !N.init_rank || !N.destroy || !N.allreduce || !N.send || !N.recv || !N.group_start || !N.group_end) {
        jfem_set_error("libnccl is missing required symbols");
        return JFEM_ENCCL;
    }
    return JFEM_OK;
}

#define JFEM_NCCL(call)                                                                                  \
    do {                                                                                                 \
        int r__ = (call);                                                                                \
        if (r__ != 0) {                                                                                  \
            jfem_set_error("NCCL error %d at %s:%d: %s", r__, __FILE__, __LINE__, N.errstr ? N.errstr(r__) : "?"); \
            return JFEM_ENCCL;                                                                           \
        }                                                                                                \
    } while (0)

__global__ void halo_pack_kernel(long long n3, const int32_t *__restrict__ nodes, const double *__restrict__ x, double *__restrict__ buf) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    const long long q = i / 3;
    buf[i] = x[3LL * nodes[q] + (i - 3 * q)];
}

__global__ void halo_unpack_kernel(long long n3, const int32_t *__restrict__ nodes, const double *__restrict__ buf, double *__restrict__ x) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    const long long q = i / 3;
    x[3LL * nodes[q] + (i - 3 * q)] = buf[i];
}

int comm_allreduce_sum(jfem_handle *h, double *buf, int count) {
    JFEM_NCCL(N.allreduce(buf, buf, (size_t)count, NCCL_DOUBLE, NCCL_SUM, h->comm, h->stream));
    return JFEM_OK;
}

// forward halo: owned interface values of x -> neighbours' ghost slots
int halo_exchange(jfem_handle *h, double *x) {
    if (h->n_ranks <= 1 || h->nb_rank.empty()) return JFEM_OK;
    const int nnb = (int)h->nb_rank.size();
    const long long ns3 = 3 * h->send_ptr[nnb], nr3 = 3 * h->recv_ptr[nnb];
    if (ns3) halo_pack_kernel<<<(unsigned)((ns3 + 255) / 256), 256, 0, h->stream>>>(ns3, h->send_nodes.p, x, h->send_buf.p);
    JFEM_NCCL(N.group_start());
    for (int i = 0; i < nnb; i++) {
        const long long sc = 3 * (h->send_ptr[i + 1] - h->send_ptr[i]), rc = 3 * (h->recv_ptr[i + 1] - h->recv_ptr[i]);
        if (sc) JFEM_NCCL(N.send(h->send_buf.p + 3 * h->send_ptr[i], (size_t)sc, NCCL_DOUBLE, h->nb_rank[i], h->comm, h->stream));
        if (rc) JFEM_NCCL(N.recv(h->recv_buf.p + 3 * h->recv_ptr[i], (size_t)rc, NCCL_DOUBLE, h->nb_rank[i], h->comm, h->stream));
    }
    JFEM_NCCL(N.group_end());
    if (nr3) halo_unpack_kernel<<<(unsigned)((nr3 + 255) / 256), 256, 0, h->stream>>>(nr3, h->recv_nodes.p, h->recv_buf.p, x);
    JFEM_CUDA(cudaGetLastError());
    h->total_launches += 2;
    return JFEM_OK;
}

extern "C" int jfem_comm_unique_id(char *id128) {
    JFEM_TRY(nccl_load());
    nccl_uid id;
    JFEM_NCCL(N.get_uid(&id));
    memcpy(id128, id.internal, 128);
    return JFEM_OK;
}

extern "C" int jfem_comm_init(jfem_handle *h, int n_ranks, int rank, const char *id128, int64_t n_owned_nodes) {
    if (!h || n_ranks < 1 || rank < 0 || rank >= n_ranks || n_owned_nodes < 0 || n_owned_nodes > h->mesh.n_nodes) {
        jfem_set_error("jfem_comm_init: bad arguments");
        return JFEM_EINVAL;
    }
    JFEM_TRY(nccl_load());
    JFEM_CUDA(cudaSetDevice(h->device));
    nccl_uid id;
    memcpy(id.internal, id128, 128);
    JFEM_NCCL(N.init_rank(&h->comm, n_ranks, id, rank));
    h->n_ranks = n_ranks; h->rank = rank; h->n_owned_nodes = n_owned_nodes;
    return JFEM_OK;
}

extern "C" int jfem_comm_set_halo(jfem_handle *h, int nnb, const int32_t *nb_rank, const int64_t *send_ptr, const int32_t *send_nodes,
                                  const int64_t *recv_ptr, const int32_t *recv_nodes) {
    if (!h || nnb < 0) { jfem_set_error("jfem_comm_set_halo: bad arguments"); return JFEM_EINVAL; }
    JFEM_CUDA(cudaSetDevice(h->device));
    h->nb_rank.assign(nb_rank, nb_rank + nnb);
    h->send_ptr.assign(send_ptr, send_ptr + nnb + 1);
    h->recv_ptr.assign(recv_ptr, recv_ptr + nnb + 1);
    std::vector<int32_t> s(send_nodes, send_nodes + send_ptr[nnb]), r(recv_nodes, recv_nodes + recv_ptr[nnb]);
    for (auto &v : s) { v -= h->index_base; if (v < 0 || v >= h->mesh.n_nodes) { jfem_set_error("halo send node out of range"); return JFEM_EINVAL; } }
    for (auto &v : r) { v -= h->index_base; if (v < 0 || v >= h->mesh.n_nodes) { jfem_set_error("halo recv node out of range"); return JFEM_EINVAL; } }
    JFEM_TRY(h->send_nodes.upload(s));
    JFEM_TRY(h->recv_nodes.upload(r));
    JFEM_TRY(h->send_buf.alloc(3 * s.size() + 1));
    JFEM_TRY(h->recv_buf.alloc(3 * r.size() + 1));
    return JFEM_OK;
}

extern "C" int jfem_comm_destroy(jfem_handle *h) {
    if (h && h->comm) { N.destroy(h->comm); h->comm = nullptr; h->n_ranks = 1; }
    return JFEM_OK;
}

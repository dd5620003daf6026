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
#include <string.h>

#include "handle.h"

#define AR_WORDS 8    // all-reduce mailbox slot: sequence word + up to 7 values

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

// ---- peer-to-peer halo over NVLink (CUDA IPC): every rank owns a landing buffer (2 parity halves) that its neighbours
// write into directly from their pack kernel, followed by a system-scope fence and a sequence flag; the receiver's
// unpack kernel spins on the flags, then copies into the ghost segment of x.  No host involvement, two tiny kernels per
// exchange (NCCL send/recv costs ~20-40 us of launch + protocol latency per exchange, which dominates a 50 us matvec).
struct P2PArgs {
    int n_nb;
    double *peer_land[8];        // neighbour's landing buffer (mapped), already offset to MY segment and parity half
    unsigned long long *peer_flag[8];   // neighbour's flag word for me
    long long send_off[9];       // node offsets of the per-neighbour send segments
    long long recv_off[9];
    const unsigned long long *my_flag[8];
};

__global__ void halo_push_kernel(P2PArgs a, const int32_t *__restrict__ nodes, const double *__restrict__ x, unsigned long long seq,
                                 unsigned int *ticket) {
    const long long n3 = 3 * a.send_off[a.n_nb];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (long long)gridDim.x * blockDim.x) {
        const long long q = i / 3;
        int nb = 0;
        while (q >= a.send_off[nb + 1]) nb++;
        a.peer_land[nb][i - 3 * a.send_off[nb]] = x[3LL * nodes[q] + (i - 3 * q)];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();   // one system fence per block (cumulative over the barrier), not one per warp
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {            // last block: all remote stores of this rank are fenced -> publish
            *ticket = 0;
            __threadfence_system();
            for (int nb = 0; nb < a.n_nb; nb++) *((volatile unsigned long long *)a.peer_flag[nb]) = seq;
        }
    }
}

__global__ void halo_pull_kernel(P2PArgs a, const double *__restrict__ land, const int32_t *__restrict__ nodes, double *__restrict__ x,
                                 unsigned long long seq) {
    if (threadIdx.x < a.n_nb) {
        const volatile unsigned long long *f = a.my_flag[threadIdx.x];
        while (*f < seq) { }
        __threadfence_system();
    }
    __syncthreads();
    const long long n3 = 3 * a.recv_off[a.n_nb];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n3; i += (long long)gridDim.x * blockDim.x) {
        const long long q = i / 3;
        x[3LL * nodes[q] + (i - 3 * q)] = __ldcv(land + i);
    }
}

// the exchange the next patch kernel performs itself (sequence number already advanced by halo_exchange)
void halo_fill_fused(jfem_handle *h, HaloFused &f) {
    const int nnb = (int)h->nb_rank.size();
    const int par = (int)(h->p2p_seq & 1);
    f.n_nb = nnb;
    for (int i = 0; i < nnb; i++) {
        f.peer_land[i] = h->p2p_peer_land[i] + (size_t)par * h->p2p_peer_half[i] + 3 * h->p2p_peer_off[i];
        f.peer_flag[i] = h->p2p_peer_flag[i];
        f.my_flag[i] = h->p2p_flags.p + h->nb_rank[i];
    }
    for (int i = 0; i <= nnb; i++) f.send_off[i] = h->send_ptr[i];
    f.send_nodes = h->send_nodes.p;
    f.land = h->p2p_land.p + (size_t)par * h->p2p_half;
    f.seq = h->p2p_seq;
    f.ticket = h->p2p_ticket.p;
}

static int halo_exchange_p2p(jfem_handle *h, double *x, bool fused_ok) {
    const int nnb = (int)h->nb_rank.size();
    h->p2p_seq++;
    if (fused_ok && h->fused_halo && h->recv_contiguous && ws_halo_capable(h)) {   // the next op_apply(x) pushes / waits / reads the landing buffer itself
        h->halo_armed = true; h->halo_x = x;
        return JFEM_OK;
    }
    const int par = (int)(h->p2p_seq & 1);
    P2PArgs a;
    a.n_nb = nnb;
    for (int i = 0; i < nnb; i++) {
        a.peer_land[i] = h->p2p_peer_land[i] + (size_t)par * h->p2p_peer_half[i] + 3 * h->p2p_peer_off[i];
        a.peer_flag[i] = h->p2p_peer_flag[i];
        a.my_flag[i] = h->p2p_flags.p + h->nb_rank[i];
    }
    for (int i = 0; i <= nnb; i++) { a.send_off[i] = h->send_ptr[i]; a.recv_off[i] = h->recv_ptr[i]; }
    const long long ns3 = 3 * h->send_ptr[nnb], nr3 = 3 * h->recv_ptr[nnb];
    int gs = (int)((ns3 + 255) / 256); if (gs > 64) gs = 64; if (gs < 1) gs = 1;
    int gr = (int)((nr3 + 255) / 256); if (gr > 64) gr = 64; if (gr < 1) gr = 1;
    halo_push_kernel<<<gs, 256, 0, h->stream>>>(a, h->send_nodes.p, x, h->p2p_seq, h->p2p_ticket.p);
    halo_pull_kernel<<<gr, 256, 0, h->stream>>>(a, h->p2p_land.p + (size_t)par * h->p2p_half, h->recv_nodes.p, x, h->p2p_seq);
    JFEM_CUDA(cudaGetLastError());
    h->total_launches += 2;
    return JFEM_OK;
}

extern "C" int jfem_comm_p2p_export(jfem_handle *h, char *handles128) {
    // allocate the landing buffer + flags, return two 64-byte IPC handles (landing, flags)
    if (!h || h->n_ranks <= 1) { jfem_set_error("jfem_comm_p2p_export: communicator not initialised"); return JFEM_ESTATE; }
    JFEM_CUDA(cudaSetDevice(h->device));
    const int nnb = (int)h->nb_rank.size();
    if (nnb > 8) { jfem_set_error("p2p halo supports at most 8 neighbours"); return JFEM_EINVAL; }
    h->p2p_half = (size_t)3 * h->recv_ptr[nnb] + 8;
    JFEM_TRY(h->p2p_land.alloc(2 * h->p2p_half));
    const size_t ctrl_words = (size_t)h->n_ranks + 1 + 2 * (size_t)h->n_ranks * AR_WORDS;   // halo flags + all-reduce mailboxes
    JFEM_TRY(h->p2p_flags.alloc(ctrl_words));
    JFEM_TRY(h->p2p_ticket.alloc(1));
    JFEM_CUDA(cudaMemset(h->p2p_flags.p, 0, ctrl_words * sizeof(unsigned long long)));
    JFEM_CUDA(cudaMemset(h->p2p_ticket.p, 0, sizeof(unsigned int)));
    cudaIpcMemHandle_t hl, hf;
    JFEM_CUDA(cudaIpcGetMemHandle(&hl, h->p2p_land.p));
    JFEM_CUDA(cudaIpcGetMemHandle(&hf, h->p2p_flags.p));
    memcpy(handles128, &hl, 64);
    memcpy(handles128 + 64, &hf, 64);
    return JFEM_OK;
}

// all_handles: n_ranks x 128 bytes (from every rank's export); recv_offsets: n_ranks x n_ranks int64 matrix where
// entry [r][s] = node offset, inside rank r's landing half, of the segment that rank s writes (-1 if none);
// halves: n_ranks int64 = p2p_half (in doubles) of every rank.
extern "C" int jfem_comm_p2p_import(jfem_handle *h, const char *all_handles, const int64_t *recv_offsets, const int64_t *halves) {
    if (!h || h->n_ranks <= 1 || !h->p2p_land.p) { jfem_set_error("jfem_comm_p2p_import: call jfem_comm_p2p_export first"); return JFEM_ESTATE; }
    JFEM_CUDA(cudaSetDevice(h->device));
    const int nnb = (int)h->nb_rank.size();
    h->p2p_peer_land.assign(nnb, nullptr); h->p2p_peer_flag.assign(nnb, nullptr);
    h->p2p_peer_off.assign(nnb, 0); h->p2p_peer_half.assign(nnb, 0);
    for (int i = 0; i < nnb; i++) {
        const int s = h->nb_rank[i];
        cudaIpcMemHandle_t hl, hf;
        memcpy(&hl, all_handles + (size_t)s * 128, 64);
        memcpy(&hf, all_handles + (size_t)s * 128 + 64, 64);
        void *pl = nullptr, *pf = nullptr;
        JFEM_CUDA(cudaIpcOpenMemHandle(&pl, hl, cudaIpcMemLazyEnablePeerAccess));
        h->p2p_opened.push_back(pl);
        JFEM_CUDA(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
        h->p2p_opened.push_back(pf);
        h->p2p_peer_land[i] = (double *)pl;
        h->p2p_peer_flag[i] = (unsigned long long *)pf + h->rank;
        const int64_t off = recv_offsets[(size_t)s * h->n_ranks + h->rank];
        if (off < 0) { jfem_set_error("rank %d does not expect halo data from rank %d", s, h->rank); return JFEM_EINVAL; }
        h->p2p_peer_off[i] = off;
        h->p2p_peer_half[i] = (size_t)halves[s];
    }
    // control buffers of all ranks (all-reduce mailboxes)
    h->p2p_ctrl.assign(h->n_ranks, nullptr);
    for (int s = 0; s < h->n_ranks; s++) {
        if (s == h->rank) { h->p2p_ctrl[s] = h->p2p_flags.p; continue; }
        bool found = false;
        for (int i = 0; i < nnb; i++)
            if (h->nb_rank[i] == s) { h->p2p_ctrl[s] = h->p2p_peer_flag[i] - h->rank; found = true; }
        if (found) continue;
        cudaIpcMemHandle_t hf;
        memcpy(&hf, all_handles + (size_t)s * 128 + 64, 64);
        void *pf = nullptr;
        JFEM_CUDA(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
        h->p2p_opened.push_back(pf);
        h->p2p_ctrl[s] = (unsigned long long *)pf;
    }
    h->p2p_ready = true;
    return JFEM_OK;
}

// ---- all-reduce of a few doubles through the same IPC-mapped control buffers: every rank writes its partial sums into
// its slot of every rank's mailbox (values, system fence, sequence word), then sums the n_ranks slots of its own
// mailbox in RANK ORDER (deterministic, identical on all ranks).  One single-warp kernel, ~NVLink write latency,
// instead of an NCCL all-reduce launch (~15-25 us for 8 bytes).

struct P2PAllArgs {
    int n_ranks, rank;
    unsigned long long *ctrl[8];   // control buffer of every rank (own included)
    long long mail_off;            // offset (u64 words) of the mailboxes inside a control buffer
};

__global__ void p2p_allreduce_kernel(P2PAllArgs a, double *sums, int count, unsigned long long seq) {
    const int r = threadIdx.x;
    const int par = (int)(seq & 1);
    if (r < a.n_ranks) {
        volatile unsigned long long *slot = a.ctrl[r] + a.mail_off + ((size_t)par * a.n_ranks + a.rank) * AR_WORDS;
        for (int c = 0; c < count; c++) slot[1 + c] = (unsigned long long)__double_as_longlong(sums[c]);
        __threadfence_system();
        slot[0] = seq;
        volatile unsigned long long *mine = a.ctrl[a.rank] + a.mail_off + ((size_t)par * a.n_ranks + r) * AR_WORDS;
        while (mine[0] != seq) { }
        __threadfence_system();
    }
    __syncwarp();
    if (r < count) {
        double acc = 0.0;
        for (int q = 0; q < a.n_ranks; q++) {
            volatile unsigned long long *mine = a.ctrl[a.rank] + a.mail_off + ((size_t)par * a.n_ranks + q) * AR_WORDS;
            acc += __longlong_as_double((long long)mine[1 + r]);
        }
        sums[r] = acc;
    }
}

int comm_allreduce_sum(jfem_handle *h, double *buf, int count) {
    if (h->p2p_ready && h->n_ranks <= 8 && count < AR_WORDS) {
        P2PAllArgs a;
        a.n_ranks = h->n_ranks; a.rank = h->rank; a.mail_off = h->n_ranks + 1;
        for (int r = 0; r < h->n_ranks; r++) a.ctrl[r] = h->p2p_ctrl[r];
        h->p2p_ar_seq++;
        p2p_allreduce_kernel<<<1, 32, 0, h->stream>>>(a, buf, count, h->p2p_ar_seq);
        JFEM_CUDA(cudaGetLastError());
        h->total_launches++;
        return JFEM_OK;
    }
    JFEM_NCCL(N.allreduce(buf, buf, (size_t)count, NCCL_DOUBLE, NCCL_SUM, h->comm, h->stream));
    return JFEM_OK;
}

// forward halo: owned interface values of x -> neighbours' ghost slots
int halo_exchange(jfem_handle *h, double *x, bool fused_ok) {
    if (h->n_ranks <= 1 || h->nb_rank.empty()) return JFEM_OK;
    if (h->p2p_ready) return halo_exchange_p2p(h, x, fused_ok);
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
    h->built = false;   // the patch order depends on which nodes are ghosts
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
    h->recv_contiguous = h->n_owned_nodes >= 0;
    for (size_t i = 0; i < r.size() && h->recv_contiguous; i++) h->recv_contiguous = r[i] == (int64_t)h->n_owned_nodes + (int64_t)i;
    JFEM_TRY(h->send_nodes.upload(s));
    JFEM_TRY(h->recv_nodes.upload(r));
    JFEM_TRY(h->send_buf.alloc(3 * s.size() + 1));
    JFEM_TRY(h->recv_buf.alloc(3 * r.size() + 1));
    return JFEM_OK;
}

extern "C" int jfem_comm_p2p_seq(jfem_handle *h, int64_t set_to, int64_t *seq) {
    if (!h) { jfem_set_error("null handle"); return JFEM_EINVAL; }
    if (set_to >= 0) { h->p2p_seq = (unsigned long long)set_to; h->halo_armed = false; }
    if (seq) *seq = (int64_t)h->p2p_seq;
    return JFEM_OK;
}

extern "C" int jfem_comm_destroy(jfem_handle *h) {
    if (!h) return JFEM_OK;
    // peer mappings opened by jfem_comm_p2p_import, then this rank's own landing / flag buffers
    for (void *p : h->p2p_opened) cudaIpcCloseMemHandle(p);
    h->p2p_opened.clear();
    h->p2p_peer_land.clear(); h->p2p_peer_flag.clear(); h->p2p_ctrl.clear();
    h->p2p_ready = false; h->halo_armed = false;
    h->p2p_land.release(); h->p2p_flags.release(); h->p2p_ticket.release();
    if (h->comm) { N.destroy(h->comm); h->comm = nullptr; h->n_ranks = 1; }
    return JFEM_OK;
}

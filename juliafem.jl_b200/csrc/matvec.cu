// matvec.cu -- the matrix-free element operator  y = A(x)  on patches (one thread block = one patch).
//
// Replaces compute_element_stresses_kernel! + nodal_assembly_kernel! (ext/JuliaFEMCUDAExt.jl:222-361) and the
// atomic Tet10 residual kernel of demos/gpu_assembly_tet10.jl:145-232.  Per block:
//   phase 0  gather x (and coordinates, linearisation point) of the patch's nodes into shared memory
//   phase 1  one thread per element: register-resident element contraction (elem.cuh), 3*nnpe results
//            written to a [dof][thread] staging tile (conflict-free)
//   phase 2  one thread per patch node: sum the staged contributions of the node's elements in a fixed
//            order (deterministic, no atomics); interior nodes are stored to y directly, interface nodes go to
//            one partial slot per (patch,node) which iface_reduce_kernel adds in ascending patch order.
#include <chrono>
#include <cstring>

#include "elem.cuh"
#include "handle.h"

using namespace jf;

struct PatchKArgs {
    const int32_t *pnode_ptr;
    const uint32_t *pnodes;
    const int32_t *n_iface;
    const int32_t *ipart_base;
    const uint16_t *lconn;
    const uint16_t *goff;
    const uint16_t *gslots;
    const double *coords;
    const double *x;
    const double *ulin;
    double *y;
    double *ipart;
    long long n_elems, elem_offset;
    int max_nodes, project, atomic_iface;
    int *fail;
    const int *done;
};

template <int NNPE, int CLS, int MODE, class Pt, int T>
__global__ void __launch_bounds__(T) patch_kernel(PatchKArgs a, Pt pt) {
    extern __shared__ double sm[];
    if (a.done && *a.done) return;
    constexpr int NF = Pt::NF;
    const int p = blockIdx.x, tid = threadIdx.x;
    const int nb = a.pnode_ptr[p], np = a.pnode_ptr[p + 1] - nb;
    double *stage = sm;
    double *xs = sm + 3 * NNPE * T;
    double *Xs = xs + 3 * a.max_nodes;
    double *us = Xs + 3 * a.max_nodes;

    // ---- phase 0: gather
    for (int j = tid; j < np; j += T) {
        const uint32_t w = a.pnodes[nb + j];
        const long long id = w & PN_ID_MASK;
        const double *px = a.x + 3 * id;
        xs[3 * j + 0] = __ldg(px); xs[3 * j + 1] = __ldg(px + 1); xs[3 * j + 2] = __ldg(px + 2);
        if (w & PN_NEEDX) {
            const double *pc = a.coords + 3 * id;
            Xs[3 * j + 0] = __ldg(pc); Xs[3 * j + 1] = __ldg(pc + 1); Xs[3 * j + 2] = __ldg(pc + 2);
        }
        if (NF == 2) {
            const double *pu = a.ulin + 3 * id;
            us[3 * j + 0] = __ldg(pu); us[3 * j + 1] = __ldg(pu + 1); us[3 * j + 2] = __ldg(pu + 2);
        }
    }
    __syncthreads();

    // ---- phase 1: element contraction
    const long long el = (long long)p * T + tid;
    if (el < a.n_elems) {
        int n[NNPE];
        JF_UNROLL for (int k = 0; k < NNPE; k++) n[k] = a.lconn[((size_t)p * NNPE + k) * T + tid];
        double *st = stage + tid;
        auto out = [st](int k, int c, double v) { st[(3 * k + c) * T] = v; };
        SField X{Xs, n};
        bool ok = true;
        if (CLS == CLASS_AFFINE && MODE == OP_LINEAR && NNPE == 10) {
            SField U{xs, n};
            tet10_affine_linear(pt.la, pt.mu, U, X, out);
        } else {
            SField F[NF];
            F[0].base = xs; F[0].n = n;
            if (NF == 2) { F[NF - 1].base = us; F[NF - 1].n = n; }
            if (NNPE == 10) ok = tet10_general(pt, a.elem_offset + el, F, X, out);
            else if (NNPE == 8) ok = hex8_general(pt, a.elem_offset + el, F, X, out);
            else ok = tet4_general(pt, a.elem_offset + el, F, X, out);
        }
        if (!ok) atomicOr(a.fail, 1);
    }
    __syncthreads();

    // ---- phase 2: per-node ordered reduction
    const uint16_t *go = a.goff + nb + p;
    const uint16_t *gs = a.gslots + (size_t)p * T * NNPE;
    const int nif = a.n_iface[p];
    for (int j = tid; j < np; j += T) {
        const int q0 = go[j], q1 = go[j + 1];
        double s0 = 0, s1 = 0, s2 = 0;
        for (int q = q0; q < q1; q++) {
            const int sl = gs[q];
            s0 += stage[sl]; s1 += stage[sl + T]; s2 += stage[sl + 2 * T];
        }
        const uint32_t w = a.pnodes[nb + j];
        if (a.project) {
            if (w & (1u << PN_FIXSHIFT)) s0 = 0.0;
            if (w & (2u << PN_FIXSHIFT)) s1 = 0.0;
            if (w & (4u << PN_FIXSHIFT)) s2 = 0.0;
        }
        if (j < nif) {
            if (a.atomic_iface) {
                double *py = a.y + 3 * (long long)(w & PN_ID_MASK);
                atomicAdd(py, s0); atomicAdd(py + 1, s1); atomicAdd(py + 2, s2);
            } else {
                double *pp = a.ipart + 3 * ((long long)a.ipart_base[p] + j);
                pp[0] = s0; pp[1] = s1; pp[2] = s2;
            }
        } else {
            double *py = a.y + 3 * (long long)(w & PN_ID_MASK);
            py[0] = s0; py[1] = s1; py[2] = s2;
        }
    }
}

// y[interface node] = sum of its partial slots, ascending (set, patch) order.  3 threads per node.
__global__ void iface_reduce_kernel(const uint32_t *__restrict__ inodes, const int32_t *__restrict__ iptr,
                                    const int32_t *__restrict__ islots, const double *__restrict__ ipart,
                                    double *__restrict__ y, long long n3, const int *done) {
    if (done && *done) return;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    const int node = (int)(i / 3), c = (int)(i - 3LL * node);
    const uint32_t w = inodes[node];
    double s = 0;
    for (int q = iptr[node]; q < iptr[node + 1]; q++) s += ipart[3LL * islots[q] + c];
    y[3LL * (w & PN_ID_MASK) + c] = s;
}

__global__ void iface_zero_kernel(const uint32_t *__restrict__ inodes, double *__restrict__ y, long long n3, const int *done) {
    if (done && *done) return;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    const int node = (int)(i / 3), c = (int)(i - 3LL * node);
    y[3LL * (inodes[node] & PN_ID_MASK) + c] = 0.0;
}

// ------------------------------------------------------------------------------------------------ build

static void apply_fixed_words(jfem_handle *h, std::vector<uint32_t> &words) {
    for (auto &w : words) {
        int64_t id = w & PN_ID_MASK;
        w &= ~(7u << PN_FIXSHIFT);
        for (int c = 0; c < 3; c++)
            if (h->mesh.fixed[3 * id + c]) w |= (1u << (PN_FIXSHIFT + c));
    }
}

int upload_fixed(jfem_handle *h) {   // (re)upload everything that embeds the Dirichlet mask
    if (!h->built) return JFEM_OK;
    for (int c = 0; c < N_CLASSES; c++) {
        if (h->hsets[c].n_elems == 0) continue;
        std::vector<uint32_t> w = h->hsets[c].pnodes;
        apply_fixed_words(h, w);
        JFEM_TRY(h->dsets[c].pnodes.upload(w));
    }
    std::vector<uint32_t> w = h->hif.inodes;
    apply_fixed_words(h, w);
    JFEM_TRY(h->inodes.upload(w));
    JFEM_TRY(h->fixed.upload(h->mesh.fixed));
    return JFEM_OK;
}

int ensure_built(jfem_handle *h) {
    if (h->built) return JFEM_OK;
    JFEM_CUDA(cudaSetDevice(h->device));
    auto t0 = std::chrono::steady_clock::now();
    classify_elements(h->mesh, h->affine);
    JFEM_TRY(build_patch_sets(h->mesh, h->patch_elems, h->affine, h->hsets, h->hif));
    h->setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int64_t off = 0;
    std::vector<int64_t> e2i(h->mesh.n_elems);
    for (int c = 0; c < N_CLASSES; c++) {
        PatchSetHost &S = h->hsets[c];
        PatchSetDev &D = h->dsets[c];
        D.release();
        D.cls = c; D.nnpe = S.nnpe; D.EP = S.EP; D.n_patches = S.n_patches; D.max_nodes = S.max_nodes;
        D.n_elems = S.n_elems; D.elem_offset = off;
        for (int64_t i = 0; i < S.n_elems; i++) e2i[S.elem_perm[i]] = off + i;
        off += S.n_elems;
        if (S.n_elems == 0) continue;
        JFEM_TRY(D.pnode_ptr.upload(S.pnode_ptr));
        JFEM_TRY(D.n_iface.upload(S.n_iface));
        JFEM_TRY(D.ipart_base.upload(S.ipart_base));
        JFEM_TRY(D.lconn.upload(S.lconn));
        JFEM_TRY(D.goff.upload(S.goff));
        JFEM_TRY(D.gslots.upload(S.gslots));
        // host copies of the big tables are no longer needed
        std::vector<uint16_t>().swap(S.lconn);
        std::vector<uint16_t>().swap(S.gslots);
        std::vector<uint16_t>().swap(S.goff);
    }
    JFEM_TRY(h->e2i.upload(e2i));
    JFEM_TRY(h->iptr.upload(h->hif.iptr));
    JFEM_TRY(h->islots.upload(h->hif.islots));
    JFEM_TRY(h->ipart.alloc((size_t)3 * h->hif.n_partials));
    JFEM_TRY(h->coords.upload(h->mesh.coords));
    JFEM_TRY(h->dflags.alloc(4));
    JFEM_CUDA(cudaMemset(h->dflags.p, 0, 4 * sizeof(int)));
    h->built = true;
    JFEM_TRY(upload_fixed(h));
    if (h->mat_kind == JFEM_MAT_PERFECT_PLASTICITY && h->st_old.n == 0) {
        size_t n = (size_t)JFEM_NSTATE * h->ngp() * h->mesh.n_elems;
        JFEM_TRY(h->st_old.alloc(n));
        JFEM_TRY(h->st_new.alloc(n));
        JFEM_CUDA(cudaMemset(h->st_old.p, 0, n * sizeof(double)));
        JFEM_CUDA(cudaMemset(h->st_new.p, 0, n * sizeof(double)));
    }
    return JFEM_OK;
}

// ------------------------------------------------------------------------------------------------ launch

template <int NNPE, int CLS, int MODE, class Pt, int T>
static int launch_set(jfem_handle *h, const PatchSetDev &D, PatchKArgs a, const Pt &pt) {
    constexpr int NF = Pt::NF;
    size_t smem = sizeof(double) * (3 * NNPE * T + (size_t)3 * D.max_nodes * (NF == 2 ? 3 : 2));
    auto kern = patch_kernel<NNPE, CLS, MODE, Pt, T>;
    static size_t configured = 0;
    if (smem > configured) {
        JFEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    if (smem > 227 * 1024) { jfem_set_error("patch needs %zu bytes of shared memory; lower patch_elems", smem); return JFEM_EINVAL; }
    kern<<<D.n_patches, T, smem, h->stream>>>(a, pt);
    JFEM_CUDA(cudaGetLastError());
    h->matvec_launches++;
    return JFEM_OK;
}

template <int NNPE, int CLS, int T>
static int dispatch_mode(jfem_handle *h, const PatchSetDev &D, PatchKArgs a, int mode) {
    double la = h->mat[0] * h->mat[1] / ((1.0 + h->mat[1]) * (1.0 - 2.0 * h->mat[1]));   // linear_elastic.jl:82
    double mu = h->mat[0] / (2.0 * (1.0 + h->mat[1]));                                    // :97
    const long long n_gp = (long long)h->mesh.n_elems * h->ngp();
    if (h->mat_kind == JFEM_MAT_LINEAR_ELASTIC || (mode == OP_LINEAR)) {
        // small-strain linear elasticity: K.x = f_int(x), tangent == K
        PtLinear pt; pt.la = la; pt.mu = mu;
        return launch_set<NNPE, CLS, OP_LINEAR, PtLinear, T>(h, D, a, pt);
    }
    if (h->mat_kind == JFEM_MAT_NEO_HOOKEAN) {
        if (mode == OP_RESIDUAL) { PtNHResidual pt; pt.la = la; pt.mu = mu; return launch_set<NNPE, CLASS_GENERAL, OP_RESIDUAL, PtNHResidual, T>(h, D, a, pt); }
        PtNHTangent pt; pt.la = la; pt.mu = mu;
        return launch_set<NNPE, CLASS_GENERAL, OP_TANGENT, PtNHTangent, T>(h, D, a, pt);
    }
    if (h->mat_kind == JFEM_MAT_PERFECT_PLASTICITY) {
        if (mode == OP_RESIDUAL) {
            PtPPResidual pt; pt.la = la; pt.mu = mu; pt.sy = h->mat[2]; pt.H = h->mat[3];
            pt.st_old = h->st_old.p; pt.st_new = h->st_new.p; pt.n_gp = n_gp;
            return launch_set<NNPE, CLASS_GENERAL, OP_RESIDUAL, PtPPResidual, T>(h, D, a, pt);
        }
        PtPPTangent pt; pt.la = la; pt.mu = mu; pt.sy = h->mat[2]; pt.H = h->mat[3];
        pt.st_old = h->st_old.p; pt.n_gp = n_gp;
        return launch_set<NNPE, CLASS_GENERAL, OP_TANGENT, PtPPTangent, T>(h, D, a, pt);
    }
    jfem_set_error("material not set");
    return JFEM_ESTATE;
}

template <int NNPE, int T>
static int dispatch_class(jfem_handle *h, const PatchSetDev &D, PatchKArgs a, int mode) {
    if (D.cls == CLASS_AFFINE) return dispatch_mode<NNPE, CLASS_AFFINE, T>(h, D, a, mode);
    return dispatch_mode<NNPE, CLASS_GENERAL, T>(h, D, a, mode);
}

template <int NNPE>
static int dispatch_threads(jfem_handle *h, const PatchSetDev &D, PatchKArgs a, int mode) {
    switch (D.EP) {
        case 128: return dispatch_class<NNPE, 128>(h, D, a, mode);
        case 256: return dispatch_class<NNPE, 256>(h, D, a, mode);
        case 512: return dispatch_class<NNPE, 512>(h, D, a, mode);
    }
    jfem_set_error("patch_elems must be 128, 256 or 512");
    return JFEM_EINVAL;
}

// y = A(x): mode OP_LINEAR (K x), OP_RESIDUAL (f_int(x)), OP_TANGENT (K(ulin) x)
int op_apply(jfem_handle *h, int mode, const double *x, double *y, int flags, const int *done) {
    JFEM_TRY(ensure_built(h));
    if (h->mat_kind < 0) { jfem_set_error("jfem_set_material has not been called"); return JFEM_ESTATE; }
    if (mode == OP_TANGENT && h->mat_kind != JFEM_MAT_LINEAR_ELASTIC && !h->has_lin) {
        jfem_set_error("tangent operator needs jfem_set_linearization"); return JFEM_ESTATE;
    }
    h->matvec_launches = 0;
    const long long n3 = 3LL * (long long)h->hif.inodes.size();
    const int atomic_iface = h->deterministic ? 0 : 1;
    if (atomic_iface && n3) {
        iface_zero_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, h->stream>>>(h->inodes.p, y, n3, done);
        h->matvec_launches++;
    }
    for (int c = 0; c < N_CLASSES; c++) {
        const PatchSetDev &D = h->dsets[c];
        if (D.n_elems == 0) continue;
        PatchKArgs a;
        a.pnode_ptr = D.pnode_ptr.p; a.pnodes = D.pnodes.p; a.n_iface = D.n_iface.p; a.ipart_base = D.ipart_base.p;
        a.lconn = D.lconn.p; a.goff = D.goff.p; a.gslots = D.gslots.p;
        a.coords = h->coords.p; a.x = x; a.ulin = h->ulin.p; a.y = y; a.ipart = h->ipart.p;
        a.n_elems = D.n_elems; a.elem_offset = D.elem_offset; a.max_nodes = D.max_nodes;
        a.project = (flags & JFEM_PROJECT) ? 1 : 0; a.atomic_iface = atomic_iface;
        a.fail = h->dflags.p; a.done = done;
        int rc;
        switch (h->mesh.nnpe) {
            case 10: rc = dispatch_threads<10>(h, D, a, mode); break;
            case 8: rc = dispatch_threads<8>(h, D, a, mode); break;
            default: rc = dispatch_threads<4>(h, D, a, mode); break;
        }
        JFEM_TRY(rc);
    }
    if (!atomic_iface && n3) {
        iface_reduce_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, h->stream>>>(h->inodes.p, h->iptr.p, h->islots.p, h->ipart.p, y, n3, done);
        JFEM_CUDA(cudaGetLastError());
        h->matvec_launches++;
    }
    h->total_launches += h->matvec_launches;
    return JFEM_OK;
}

// assemble.cu -- element matrices, global CSR pattern + coloured scatter, SpMV.
//
// Replaces assemble_element! -> add!(SparseMatrixCOO) -> sparse() (src/problems_elasticity.jl:203-451,
// src/sparse/sparse.jl:53-55,121-132).  The pattern is exactly the reference's: every (dof_i, dof_j) pair of every
// element is stored, zeros included, rows = 3*(node-1)+c, column indices ascending.  Because the pattern is the
// node-adjacency graph blown up by 3x3, it is built at node level on the host (sort + unique per node) and expanded
// arithmetically:  rowptr[3n+c] = 9*adjptr[n] + 3*c*deg(n),  colind = 3*adj + d.
// Values: one thread per (element, column j) applies the tangent functor of elem.cuh to the unit vector e_j, which
// yields column j of Ke = Km(+Kg); elements are processed colour by colour (greedy colouring, the scheme of
// src/preprocess.jl:331-398) so that the += into vals needs no atomics and the summation order is fixed.
#include <algorithm>
#include <chrono>
#include <exception>
#include <thread>

#include "asm_elem.cuh"
#include "handle.h"

using namespace jf;

void fill_material(const jfem_handle *h, jf::MatBase &m);

// global-memory field accessor; base == nullptr means "unit vector e_(kj,cj)"
struct AField {
    const double *base;
    const int *n;
    int kj, cj;
    __device__ __forceinline__ double operator()(int k, int c) const {
        return base ? __ldg(base + 3 * (long long)n[k] + c) : ((k == kj && c == cj) ? 1.0 : 0.0);
    }
};

template <int NNPE, class Pt, class OUT>
__device__ __forceinline__ bool elem_dispatch(const Pt &pt, long long ei, const AField (&F)[Pt::NF], const AField &X, OUT &&out) {
    if (NNPE == 10) return tet10_general(pt, ei, F, X, out);
    if (NNPE == 8) return hex8_general(pt, ei, F, X, out);
    return tet4_general(pt, ei, F, X, out);
}

struct AsmArgs {
    const int32_t *conn;       // caller order, 0-based
    const int32_t *elems;      // element list of this launch (nullptr: e0 + i)
    long long e0, ne;
    const double *coords, *u;
    const long long *e2i;
    const long long *adjptr;
    const uint16_t *eblk;
    double *vals;              // CSR values (scatter) ...
    double *Ke;                // ... or dense per-element output (column-major), one of the two
    int *fail;
};

template <int NNPE, class Pt>
__global__ void __launch_bounds__(128) elem_columns_kernel(AsmArgs a, Pt pt) {
    constexpr int ND = 3 * NNPE, NF = Pt::NF;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.ne * ND) return;
    const long long i = idx / ND;
    const int j = (int)(idx - i * ND), kj = j / 3, cj = j - 3 * kj;
    const long long e = a.elems ? a.elems[a.e0 + i] : a.e0 + i;
    int n[NNPE];
    JF_UNROLL for (int k = 0; k < NNPE; k++) n[k] = a.conn[e * NNPE + k];
    AField X{a.coords, n, 0, 0};
    AField F[NF];
    F[0] = AField{nullptr, n, kj, cj};
    if (NF == 2) F[NF - 1] = AField{a.u, n, 0, 0};
    bool ok;
    if (a.vals) {
        double *vals = a.vals;
        const long long *adjptr = a.adjptr;
        const uint16_t *blk = a.eblk + e * NNPE * NNPE;
        auto out = [=](int k, double v0, double v1, double v2) {
            const long long ap = adjptr[n[k]], deg = adjptr[n[k] + 1] - ap;
            double *d = vals + 9 * ap + 3 * blk[k * NNPE + kj] + cj;
            d[0] += v0; d[3 * deg] += v1; d[6 * deg] += v2;
        };
        ok = elem_dispatch<NNPE>(pt, a.e2i[e], F, X, out);
    } else {
        double *Ke = a.Ke + (i * ND + j) * ND;
        auto out = [=](int k, double v0, double v1, double v2) { Ke[3 * k] = v0; Ke[3 * k + 1] = v1; Ke[3 * k + 2] = v2; };
        ok = elem_dispatch<NNPE>(pt, a.e2i[e], F, X, out);
    }
    if (!ok) atomicOr(a.fail, 1);
}

// One warp per element (default for the assembled path): lanes 0..NGP-1 compute the geometry of one Gauss point each
// (grad N_k, w, and grad u for the two-field tangent functors) into shared memory, then lane j < 3*NNPE builds column j
// of Ke from it (asm_elem.cuh) and adds its 3*NNPE entries to the CSR values (or stores them to the dense output).
// Four elements per block; the geometry is computed once per element instead of once per column.
template <int NNPE, class Pt>
__global__ void __launch_bounds__(128) elem_warp_kernel(AsmArgs a, Pt pt0) {
    constexpr int ND = 3 * NNPE, NF = Pt::NF, NGP = ElemRule<NNPE>::NGP, WPB = 4;
    __shared__ double s_gN[WPB][NGP * 3 * NNPE];
    __shared__ double s_Gu[WPB][NGP * 9];
    __shared__ double s_w[WPB][NGP];
    __shared__ int s_n[WPB][NNPE];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long i = (long long)blockIdx.x * WPB + warp;
    if (i >= a.ne) return;   // whole warps leave together; only __syncwarp below
    const long long e = a.elems ? a.elems[a.e0 + i] : a.e0 + i;
    if (lane < NNPE) s_n[warp][lane] = a.conn[e * NNPE + lane];
    __syncwarp();
    const int *n = s_n[warp];
    if (lane < NGP) {
        AField X{a.coords, n, 0, 0};
        double *gN = &s_gN[warp][lane * 3 * NNPE];
        s_w[warp][lane] = gp_geometry<NNPE>(lane, X, gN);
        if (NF == 2) {
            AField U{a.u, n, 0, 0};
            gp_grad<NNPE>(U, gN, &s_Gu[warp][lane * 9]);
        }
    }
    __syncwarp();
    if (lane >= ND) return;
    const int l = lane / 3, cj = lane - 3 * l;
    const long long ei = a.e2i[e];
    Pt pt = pt0;
    pt.load(ei);
    double acc[NNPE][3];
    JF_UNROLL for (int k = 0; k < NNPE; k++) { acc[k][0] = 0.0; acc[k][1] = 0.0; acc[k][2] = 0.0; }
    bool ok = true;
#pragma unroll 1
    for (int g = 0; g < NGP; g++)
        ok &= gp_column<NNPE>(pt, ei * NGP + g, &s_gN[warp][g * 3 * NNPE], s_w[warp][g], &s_Gu[warp][g * 9], l, cj, acc);
    if (a.vals) {
        // read-modify-write of the 3*NNPE entries of this column, in two batches: all loads of a batch are issued before
        // its first store (the compiler may not move a load of vals across a store to vals, so a plain `+=` per entry would
        // serialise 3*NNPE global round trips per lane).  Elements of one colour share no node: no other warp touches these entries.
        const uint16_t *blk = a.eblk + e * NNPE * NNPE;
        constexpr int HALF = (NNPE + 1) / 2;
        JF_UNROLL for (int k0 = 0; k0 < NNPE; k0 += HALF) {
            double *d[HALF];
            long long s3[HALF];
            double o[HALF][3];
            JF_UNROLL for (int q = 0; q < HALF; q++) if (k0 + q < NNPE) {
                const int k = k0 + q;
                const long long ap = a.adjptr[n[k]], deg = a.adjptr[n[k] + 1] - ap;
                d[q] = a.vals + 9 * ap + 3 * blk[k * NNPE + l] + cj;
                s3[q] = 3 * deg;
                o[q][0] = d[q][0]; o[q][1] = d[q][s3[q]]; o[q][2] = d[q][2 * s3[q]];
            }
            JF_UNROLL for (int q = 0; q < HALF; q++) if (k0 + q < NNPE) {
                const int k = k0 + q;
                d[q][0] = o[q][0] + acc[k][0]; d[q][s3[q]] = o[q][1] + acc[k][1]; d[q][2 * s3[q]] = o[q][2] + acc[k][2];
            }
        }
    } else {
        double *Ke = a.Ke + (i * ND + lane) * ND;
        JF_UNROLL for (int k = 0; k < NNPE; k++) { Ke[3 * k] = acc[k][0]; Ke[3 * k + 1] = acc[k][1]; Ke[3 * k + 2] = acc[k][2]; }
    }
    if (!ok) atomicOr(a.fail, 1);
}

template <int NNPE, class Pt>
__global__ void __launch_bounds__(128) elem_fint_kernel(AsmArgs a, Pt pt, double *fe) {
    constexpr int ND = 3 * NNPE;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.ne) return;
    const long long e = a.e0 + i;
    int n[NNPE];
    JF_UNROLL for (int k = 0; k < NNPE; k++) n[k] = a.conn[e * NNPE + k];
    AField X{a.coords, n, 0, 0};
    AField F[1] = {AField{a.u, n, 0, 0}};
    double *o = fe + i * ND;
    auto out = [=](int k, double v0, double v1, double v2) { o[3 * k] = v0; o[3 * k + 1] = v1; o[3 * k + 2] = v2; };
    bool ok = elem_dispatch<NNPE>(pt, a.e2i[e], F, X, out);
    if (!ok) atomicOr(a.fail, 1);
}

__global__ void expand_pattern_kernel(long long n_nodes, const long long *__restrict__ adjptr, const int32_t *__restrict__ adj,
                                      long long *__restrict__ rowptr, int32_t *__restrict__ colind) {
    long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (a > n_nodes) return;
    if (a == n_nodes) { rowptr[3 * n_nodes] = 9 * adjptr[n_nodes]; return; }
    const long long ap = adjptr[a], deg = adjptr[a + 1] - ap;
    for (int c = 0; c < 3; c++) {
        const long long r0 = 9 * ap + 3 * c * deg;
        rowptr[3 * a + c] = r0;
        for (long long q = 0; q < deg; q++) {
            const int32_t m = adj[ap + q];
            colind[r0 + 3 * q] = 3 * m; colind[r0 + 3 * q + 1] = 3 * m + 1; colind[r0 + 3 * q + 2] = 3 * m + 2;
        }
    }
}

// K <- (K + K')/2  (src/solvers.jl:289-292); one thread per stored entry above the diagonal
__global__ void symmetrise_kernel(long long n_rows, const long long *__restrict__ rowptr, const int32_t *__restrict__ colind, double *vals) {
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    for (long long p = rowptr[r]; p < rowptr[r + 1]; p++) {
        const long long c = colind[p];
        if (c <= r) continue;
        long long lo = rowptr[c], hi = rowptr[c + 1];
        while (lo < hi) { long long m = (lo + hi) >> 1; if (colind[m] < r) lo = m + 1; else hi = m; }
        const double v = 0.5 * (vals[p] + vals[lo]);
        vals[p] = v; vals[lo] = v;
    }
}

// y = K x, one warp per row, fixed lane-strided order + shuffle tree (deterministic)
__global__ void __launch_bounds__(256) spmv_kernel(long long n_rows, const long long *__restrict__ rowptr, const int32_t *__restrict__ colind,
                                                    const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
                                                    const uint8_t *__restrict__ fixed, int project, const int *done) {
    if (done && *done) return;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    double acc = 0;
    for (long long p = rowptr[row] + lane; p < rowptr[row + 1]; p += 32) acc += vals[p] * __ldg(x + colind[p]);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[row] = (project && fixed[row]) ? 0.0 : acc;
}

// ------------------------------------------------------------------------------------------------ host side

// greedy element colouring (src/preprocess.jl:331-398), elements visited in ascending id; elements of one colour share no node
int ensure_colouring(jfem_handle *h) {
    if (!h->colour_ptr.empty()) return JFEM_OK;
    std::vector<int32_t> celems;
    greedy_colouring(h->mesh, h->colour_ptr, celems);
    JFEM_TRY(h->colour_elems.upload(celems));
    if (h->dconn.n == 0) JFEM_TRY(h->dconn.upload(h->mesh.conn));
    return JFEM_OK;
}

int csr_build(jfem_handle *h) {
    if (h->csr_built) return JFEM_OK;
    JFEM_TRY(ensure_built(h));
    auto t0 = std::chrono::steady_clock::now();
    const int64_t nn = h->mesh.n_nodes;
    std::vector<uint16_t> eblk;
    {   // the (inherently sequential) greedy colouring runs on a thread of its own next to the OpenMP adjacency build
        std::vector<int32_t> celems;
        const bool need_colours = h->colour_ptr.empty();
        std::thread colouring_thread;
        if (need_colours) colouring_thread = std::thread([&] { greedy_colouring(h->mesh, h->colour_ptr, celems); });
        int rc;
        try {
            rc = build_node_adjacency(h->mesh, h->h_nadj_ptr, h->h_nadj, eblk);   // host: hash-unique + sort per node (patches.cpp)
        } catch (const std::exception &ex) {   // (out of memory on the host): never leave a joinable thread behind
            if (need_colours) colouring_thread.join();
            h->colour_ptr.clear();
            jfem_set_error("pattern build failed: %s", ex.what());
            return JFEM_EINVAL;
        }
        if (need_colours) colouring_thread.join();
        if (rc != JFEM_OK) { h->colour_ptr.clear(); return rc; }
        if (need_colours) JFEM_TRY(h->colour_elems.upload(celems));
        if (h->dconn.n == 0) JFEM_TRY(h->dconn.upload(h->mesh.conn));
    }
    JFEM_TRY(h->nadj_ptr.upload(h->h_nadj_ptr));
    JFEM_TRY(h->nadj.upload(h->h_nadj));
    JFEM_TRY(h->eblk.upload(eblk));
    const int64_t nnz = 9 * h->h_nadj_ptr[nn];
    JFEM_TRY(h->rowptr.alloc(3 * nn + 1));
    JFEM_TRY(h->colind.alloc(nnz));
    JFEM_TRY(h->vals.alloc(nnz));
    expand_pattern_kernel<<<(unsigned)((nn + 1 + 127) / 128), 128, 0, h->stream>>>(nn, (const long long *)h->nadj_ptr.p, h->nadj.p, (long long *)h->rowptr.p, h->colind.p);
    JFEM_CUDA(cudaGetLastError());
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    h->pattern_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    h->total_launches++;
    h->csr_built = true;
    return JFEM_OK;
}

static int ensure_dconn(jfem_handle *h) {
    if (h->dconn.n == 0) JFEM_TRY(h->dconn.upload(h->mesh.conn));
    return JFEM_OK;
}

template <int NNPE, class Pt>
static int launch_columns(jfem_handle *h, AsmArgs a, const Pt &pt) {
    const long long nthreads = a.ne * 3 * NNPE;
    if (nthreads == 0) return JFEM_OK;
    if (h->asm_warp) elem_warp_kernel<NNPE, Pt><<<(unsigned)((a.ne + 3) / 4), 128, 0, h->stream>>>(a, pt);
    else elem_columns_kernel<NNPE, Pt><<<(unsigned)((nthreads + 127) / 128), 128, 0, h->stream>>>(a, pt);
    JFEM_CUDA(cudaGetLastError());
    h->total_launches++;
    return JFEM_OK;
}

template <int NNPE>
static int columns_by_material(jfem_handle *h, AsmArgs a) {
    const long long n_gp = (long long)h->mesh.n_elems * h->ngp();
    if (h->mat_kind == JFEM_MAT_LINEAR_ELASTIC) { PtLinear pt; fill_material(h, pt); return launch_columns<NNPE>(h, a, pt); }
    if (h->mat_kind == JFEM_MAT_NEO_HOOKEAN) { PtNHTangent pt; fill_material(h, pt); return launch_columns<NNPE>(h, a, pt); }
    if (h->mat_kind == JFEM_MAT_STVK) { PtStVKTangent pt; fill_material(h, pt); pt.geo = h->geometric_stiffness ? 1 : 0; return launch_columns<NNPE>(h, a, pt); }
    if (h->mat_kind == JFEM_MAT_PERFECT_PLASTICITY) {
        PtPPTangent pt; fill_material(h, pt); pt.st_old = h->st_old.p; pt.n_gp = n_gp;
        return launch_columns<NNPE>(h, a, pt);
    }
    jfem_set_error("material not set");
    return JFEM_ESTATE;
}

static int columns_dispatch(jfem_handle *h, AsmArgs a) {
    switch (h->mesh.nnpe) {
        case 10: return columns_by_material<10>(h, a);
        case 8: return columns_by_material<8>(h, a);
        default: return columns_by_material<4>(h, a);
    }
}

static int check_fail(jfem_handle *h, const char *what) {
    int fail = 0;
    JFEM_CUDA(cudaMemcpyAsync(&fail, h->dflags.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    if (fail) {
        JFEM_CUDA(cudaMemset(h->dflags.p, 0, sizeof(int)));
        jfem_set_error("Jacobian J = sqrt(det(C)) must be positive (invalid deformation in %s)", what);
        return JFEM_EDOMAIN;
    }
    return JFEM_OK;
}

int csr_assemble(jfem_handle *h, const double *u, int symmetrise) {
    JFEM_TRY(csr_build(h));
    if (h->mat_kind != JFEM_MAT_LINEAR_ELASTIC && !u) { jfem_set_error("nonlinear material needs a displacement vector"); return JFEM_EINVAL; }
    JFEM_CUDA(cudaMemsetAsync(h->vals.p, 0, h->vals.bytes(), h->stream));
    AsmArgs a;
    a.conn = h->dconn.p; a.elems = h->colour_elems.p; a.coords = h->coords.p; a.u = u; a.e2i = (const long long *)h->e2i.p;
    a.adjptr = (const long long *)h->nadj_ptr.p; a.eblk = h->eblk.p; a.vals = h->vals.p; a.Ke = nullptr; a.fail = h->dflags.p;
    for (size_t c = 0; c + 1 < h->colour_ptr.size(); c++) {
        a.e0 = h->colour_ptr[c]; a.ne = h->colour_ptr[c + 1] - h->colour_ptr[c];
        JFEM_TRY(columns_dispatch(h, a));
    }
    if (symmetrise) {
        const long long nr = h->n_dofs();
        symmetrise_kernel<<<(unsigned)((nr + 127) / 128), 128, 0, h->stream>>>(nr, (const long long *)h->rowptr.p, h->colind.p, h->vals.p);
        JFEM_CUDA(cudaGetLastError());
        h->total_launches++;
    }
    JFEM_TRY(check_fail(h, "assembly"));
    h->vals_valid = true;
    return JFEM_OK;
}

int csr_spmv(jfem_handle *h, const double *x, double *y, int flags, const int *done) {
    if (!h->vals_valid) { jfem_set_error("jfem_assemble_csr has not been called"); return JFEM_ESTATE; }
    const long long nr = h->n_dofs();
    spmv_kernel<<<(unsigned)((nr * 32 + 255) / 256), 256, 0, h->stream>>>(nr, (const long long *)h->rowptr.p, h->colind.p, h->vals.p, x, y, h->fixed.p,
                                                                           (flags & JFEM_PROJECT) ? 1 : 0, done);
    JFEM_CUDA(cudaGetLastError());
    h->total_launches++;
    h->matvec_launches = 1;
    return JFEM_OK;
}

template <int NNPE>
static int fint_by_material(jfem_handle *h, AsmArgs a, double *fe) {
    const long long n_gp = (long long)h->mesh.n_elems * h->ngp();
    const unsigned blocks = (unsigned)((a.ne + 127) / 128);
    if (h->mat_kind == JFEM_MAT_LINEAR_ELASTIC) { PtLinear pt; fill_material(h, pt); elem_fint_kernel<NNPE><<<blocks, 128, 0, h->stream>>>(a, pt, fe); }
    else if (h->mat_kind == JFEM_MAT_NEO_HOOKEAN) { PtNHResidual pt; fill_material(h, pt); elem_fint_kernel<NNPE><<<blocks, 128, 0, h->stream>>>(a, pt, fe); }
    else if (h->mat_kind == JFEM_MAT_STVK) { PtStVKResidual pt; fill_material(h, pt); pt.geo = 0; elem_fint_kernel<NNPE><<<blocks, 128, 0, h->stream>>>(a, pt, fe); }
    else {
        PtPPResidual pt; fill_material(h, pt); pt.st_old = h->st_old.p; pt.st_new = nullptr; pt.n_gp = n_gp;
        elem_fint_kernel<NNPE><<<blocks, 128, 0, h->stream>>>(a, pt, fe);
    }
    JFEM_CUDA(cudaGetLastError());
    h->total_launches++;
    return JFEM_OK;
}

// dense element matrices of elements [e0, e0+ne) in caller order (parity checks of Ke entries)
int element_matrices(jfem_handle *h, const double *u, int64_t e0, int64_t ne, double *Ke, double *fe) {
    JFEM_TRY(ensure_built(h));
    JFEM_TRY(ensure_dconn(h));
    if (h->mat_kind < 0) { jfem_set_error("jfem_set_material has not been called"); return JFEM_ESTATE; }
    if (h->mat_kind != JFEM_MAT_LINEAR_ELASTIC && !u) { jfem_set_error("nonlinear material needs a displacement vector"); return JFEM_EINVAL; }
    AsmArgs a;
    a.conn = h->dconn.p; a.elems = nullptr; a.e0 = e0; a.ne = ne; a.coords = h->coords.p; a.u = u; a.e2i = (const long long *)h->e2i.p;
    a.adjptr = nullptr; a.eblk = nullptr; a.vals = nullptr; a.Ke = Ke; a.fail = h->dflags.p;
    if (Ke) JFEM_TRY(columns_dispatch(h, a));
    if (fe) {
        if (!u) { JFEM_CUDA(cudaMemsetAsync(fe, 0, sizeof(double) * ne * 3 * h->mesh.nnpe, h->stream)); }
        else switch (h->mesh.nnpe) {
            case 10: JFEM_TRY(fint_by_material<10>(h, a, fe)); break;
            case 8: JFEM_TRY(fint_by_material<8>(h, a, fe)); break;
            default: JFEM_TRY(fint_by_material<4>(h, a, fe)); break;
        }
    }
    return check_fail(h, "element integration");
}

// 3x3 diagonal blocks of K (or of the tangent K(u)) without assembling K: one thread per (element, node k) evaluates the
// three columns (k, 0..2) of the element matrix and keeps the rows of node k; coloured launches -> race-free, ordered.
template <int NNPE, class Pt>
__global__ void __launch_bounds__(128) elem_diag_kernel(AsmArgs a, Pt pt, double *__restrict__ D) {
    constexpr int NF = Pt::NF;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.ne * NNPE) return;
    const long long i = idx / NNPE;
    const int kj = (int)(idx - i * NNPE);
    const long long e = a.elems[a.e0 + i];
    int n[NNPE];
    JF_UNROLL for (int k = 0; k < NNPE; k++) n[k] = a.conn[e * NNPE + k];
    AField X{a.coords, n, 0, 0};
    double *d = D + 9LL * n[kj];
    for (int cj = 0; cj < 3; cj++) {
        AField F[NF];
        F[0] = AField{nullptr, n, kj, cj};
        if (NF == 2) F[NF - 1] = AField{a.u, n, 0, 0};
        auto out = [=](int k, double v0, double v1, double v2) {
            if (k == kj) { d[cj] += v0; d[3 + cj] += v1; d[6 + cj] += v2; }
        };
        elem_dispatch<NNPE>(pt, a.e2i[e], F, X, out);
    }
}

// D <- inverse of the projected block (fixed dofs: unit row/column)
__global__ void invert_blocks_kernel(long long n_nodes, const uint8_t *__restrict__ fixed, double *__restrict__ D) {
    long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_nodes) return;
    double A[3][3], B[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) A[r][c] = D[9 * n + 3 * r + c];
    for (int r = 0; r < 3; r++)
        if (fixed[3 * n + r]) { for (int c = 0; c < 3; c++) { A[r][c] = 0.0; A[c][r] = 0.0; } A[r][r] = 1.0; }
    const double tr = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
    if (tr == 0.0) { for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) B[r][c] = (r == c) ? 1.0 : 0.0; }   // orphan node
    else inv3x3(A, B);
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) D[9 * n + 3 * r + c] = B[r][c];
}

template <int NNPE, class Pt>
static int launch_diag(jfem_handle *h, AsmArgs a, const Pt &pt, double *D) {
    const long long nthreads = a.ne * NNPE;
    if (nthreads == 0) return JFEM_OK;
    elem_diag_kernel<NNPE, Pt><<<(unsigned)((nthreads + 127) / 128), 128, 0, h->stream>>>(a, pt, D);
    JFEM_CUDA(cudaGetLastError());
    h->total_launches++;
    return JFEM_OK;
}

template <int NNPE>
static int diag_by_material(jfem_handle *h, AsmArgs a, double *D, bool tangent) {
    const long long n_gp = (long long)h->mesh.n_elems * h->ngp();
    if (h->mat_kind == JFEM_MAT_LINEAR_ELASTIC || !tangent) { PtLinear pt; fill_material(h, pt); return launch_diag<NNPE>(h, a, pt, D); }
    if (h->mat_kind == JFEM_MAT_NEO_HOOKEAN) { PtNHTangent pt; fill_material(h, pt); return launch_diag<NNPE>(h, a, pt, D); }
    if (h->mat_kind == JFEM_MAT_STVK) { PtStVKTangent pt; fill_material(h, pt); pt.geo = h->geometric_stiffness ? 1 : 0; return launch_diag<NNPE>(h, a, pt, D); }
    PtPPTangent pt; fill_material(h, pt); pt.st_old = h->st_old.p; pt.n_gp = n_gp;
    return launch_diag<NNPE>(h, a, pt, D);
}

// opt-in block-Jacobi preconditioner (the reference lists a preconditioner as its first missing piece,
// docs/src/book/gpu_benchmark_milestone.md:561-565): cg_dinv = inverse 3x3 diagonal blocks of the (projected) operator
int jacobi_build(jfem_handle *h, int flags) {
    JFEM_TRY(ensure_built(h));
    JFEM_TRY(ensure_colouring(h));
    const bool tangent = (flags & JFEM_TANGENT) != 0;
    const size_t n9 = 9 * (size_t)h->mesh.n_nodes;
    if (h->cg_dinv.n != n9) JFEM_TRY(h->cg_dinv.alloc(n9));
    JFEM_CUDA(cudaMemsetAsync(h->cg_dinv.p, 0, n9 * sizeof(double), h->stream));
    AsmArgs a;
    a.conn = h->dconn.p; a.elems = h->colour_elems.p; a.coords = h->coords.p; a.u = h->ulin.p; a.e2i = (const long long *)h->e2i.p;
    a.adjptr = nullptr; a.eblk = nullptr; a.vals = nullptr; a.Ke = nullptr; a.fail = h->dflags.p;
    for (size_t c = 0; c + 1 < h->colour_ptr.size(); c++) {
        a.e0 = h->colour_ptr[c]; a.ne = h->colour_ptr[c + 1] - h->colour_ptr[c];
        int rc;
        switch (h->mesh.nnpe) {
            case 10: rc = diag_by_material<10>(h, a, h->cg_dinv.p, tangent); break;
            case 8: rc = diag_by_material<8>(h, a, h->cg_dinv.p, tangent); break;
            default: rc = diag_by_material<4>(h, a, h->cg_dinv.p, tangent); break;
        }
        JFEM_TRY(rc);
    }
    const long long nn = h->mesh.n_nodes;
    invert_blocks_kernel<<<(unsigned)((nn + 127) / 128), 128, 0, h->stream>>>(nn, h->fixed.p, h->cg_dinv.p);
    JFEM_CUDA(cudaGetLastError());
    h->total_launches++;
    return JFEM_OK;
}

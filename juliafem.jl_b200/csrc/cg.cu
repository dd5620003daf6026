// cg.cu -- conjugate gradients and the inexact Newton-Krylov driver, device resident.
//
// Arithmetic of cg_solve_matfree_gpu! (ext/JuliaFEMCUDAExt.jl:531-577) / cg_solve (src/backend/cpu.jl:221-254):
//   r = b - A x ; zero fixed dofs of r ; p = r ; rr = r.r
//   loop: Ap = A p (fixed dofs zeroed) ; alpha = rr / p.Ap ; x += alpha p ; r -= alpha Ap ; rr' = r.r ;
//         stop if sqrt(rr') < tol ; beta = rr'/rr ; p = r + beta p
// What differs from the reference is only where things live: the two dot products, alpha, beta, the
// convergence test and the iteration counter stay on the device (the reference synchronises the host twice per
// iteration, ext:561,565), the vector updates are fused into two passes, and every kernel early-exits once the
// device-side `done` flag is set so the host only polls every few iterations.  Reductions are two-level with a
// fixed grid and a fixed summation order (deterministic).  With n_ranks > 1 the partial dots are all-reduced
// over NCCL (replacing MPI.Allreduce, demos/krylov_mpi_gpu_demo.jl:231-277).
#include <math.h>

#include "handle.h"

#define RED_BLOCKS 592   // 4 x 148 SMs
#define RED_THREADS 256

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[RED_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
    if (threadIdx.x < 32) {
        s = threadIdx.x < RED_THREADS / 32 ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    __syncthreads();
    return s;   // valid in thread 0
}

// Block partials of N values -> partials[(slot0+k)*RED_BLOCKS + blockIdx]; the last block to arrive (single
// ticket) sums them in index order.  Returns true in thread 0 of that block with the totals in total[].
template <int N>
__device__ __forceinline__ bool grid_sum(const double (&v)[N], double *partials, int slot0, unsigned int *ticket, double (&total)[N]) {
    double s[N];
    for (int k = 0; k < N; k++) s[k] = block_sum(v[k]);
    __shared__ bool last;
    if (threadIdx.x == 0) {
        for (int k = 0; k < N; k++) partials[(slot0 + k) * RED_BLOCKS + blockIdx.x] = s[k];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    for (int k = 0; k < N; k++) {
        double acc = 0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS) acc += ((volatile double *)partials)[(slot0 + k) * RED_BLOCKS + i];
        total[k] = block_sum(acc);
    }
    if (threadIdx.x == 0) *ticket = 0;
    return threadIdx.x == 0;
}

enum { ST_INIT = 0, ST_ALPHA = 1, ST_UPDATE = 2, ST_ZINIT = 3, ST_Z = 4 };

// scalar stage logic, run by exactly one thread once the (global) sums are known
__device__ __forceinline__ void cg_scalar_step(CGScalars *s, int stage) {
    if (stage == ST_INIT) {
        s->rr = s->rr_new;
        s->thr = s->rel ? s->thr * sqrt(s->bnorm2) : s->thr;
        s->iters = 0;
        double rn = sqrt(s->rr);
        s->done = (s->rel ? rn <= s->thr : rn < s->thr) ? 1 : 0;   // early exit, cpu.jl:230-233
        if (s->max_iter <= 0) s->done = 1;
    } else if (stage == ST_ALPHA) {
        s->alpha = (s->pcg ? s->rz : s->rr) / s->pAp;
    } else if (stage == ST_UPDATE) {
        double rn = sqrt(s->rr_new);
        s->iters += 1;
        if (!s->pcg) s->beta = s->rr_new / s->rr;
        s->rr = s->rr_new;
        if ((s->rel ? rn <= s->thr : rn < s->thr) || s->iters >= s->max_iter) s->done = 1;
    } else if (stage == ST_ZINIT) {
        s->rz = s->rz_new;
    } else {   // ST_Z: preconditioned beta = (r.z)_new / (r.z)_old
        s->beta = s->rz_new / s->rz;
        s->rz = s->rz_new;
    }
}

__global__ void cg_scalar_kernel(CGScalars *s, int stage, const double *sums) {
    if (stage != ST_INIT && s->done) return;
    if (stage == ST_INIT) { s->rr_new = sums[0]; s->bnorm2 = sums[1]; }
    else if (stage == ST_ALPHA) s->pAp = sums[0];
    else if (stage == ST_UPDATE) s->rr_new = sums[0];
    else s->rz_new = sums[0];
    cg_scalar_step(s, stage);
}

// r = P(b - Ax), p = r, rr = r.r, bnorm2 = P(b).P(b)
__global__ void __launch_bounds__(RED_THREADS) cg_init_kernel(long long n, const double *__restrict__ b, const double *__restrict__ Ax,
                                                               const uint8_t *__restrict__ fixed, double *__restrict__ r, double *__restrict__ p,
                                                               double *partials, CGScalars *s, double *sums, int fuse) {
    double rr = 0, bb = 0;
    for (long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS) {
        bool fx = fixed[i];
        double bi = fx ? 0.0 : b[i];
        double ri = fx ? 0.0 : bi - Ax[i];
        r[i] = ri; p[i] = ri;
        rr += ri * ri; bb += bi * bi;
    }
    double v[2] = {rr, bb}, t[2];
    if (grid_sum<2>(v, partials, 0, &s->ticket[0], t)) {
        sums[0] = t[0]; sums[1] = t[1];
        if (fuse) { s->rr_new = t[0]; s->bnorm2 = t[1]; cg_scalar_step(s, ST_INIT); }
    }
}

// pAp = p.Ap over owned dofs
__global__ void __launch_bounds__(RED_THREADS) cg_dot_kernel(long long n, const double *__restrict__ a, const double *__restrict__ b,
                                                              double *partials, CGScalars *s, double *sums, int fuse) {
    if (s->done) return;
    double acc = 0;
    for (long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS) acc += a[i] * b[i];
    double v[1] = {acc}, t[1];
    if (grid_sum<1>(v, partials, 0, &s->ticket[0], t)) {
        sums[0] = t[0];
        if (fuse) { s->pAp = t[0]; cg_scalar_step(s, ST_ALPHA); }
    }
}

// x += alpha p ; r -= alpha Ap ; rr_new = r.r   (n_all: all local dofs incl. ghosts for the updates; n_own for the dot)
__global__ void __launch_bounds__(RED_THREADS) cg_update_kernel(long long n_all, long long n_own, double *__restrict__ x, double *__restrict__ r,
                                                                 const double *__restrict__ p, const double *__restrict__ Ap,
                                                                 double *partials, CGScalars *s, double *sums, int fuse) {
    if (s->done) return;
    const double alpha = s->alpha;
    double acc = 0;
    for (long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n_all; i += (long long)gridDim.x * RED_THREADS) {
        x[i] += alpha * p[i];
        double ri = r[i] - alpha * Ap[i];
        r[i] = ri;
        if (i < n_own) acc += ri * ri;
    }
    double v[1] = {acc}, t[1];
    if (grid_sum<1>(v, partials, 0, &s->ticket[0], t)) {
        sums[0] = t[0];
        if (fuse) { s->rr_new = t[0]; cg_scalar_step(s, ST_UPDATE); }
    }
}

// z = Dinv r (3x3 block per node), rz = r.z ; at init also p = z.  One thread per owned node.
__global__ void __launch_bounds__(RED_THREADS) pcg_z_kernel(long long n_nodes, const double *__restrict__ r, const double *__restrict__ dinv,
                                                             double *__restrict__ z, double *__restrict__ p, double *partials, CGScalars *s,
                                                             double *sums, int fuse, int stage) {
    if (stage == ST_Z && s->done) return;
    double acc = 0;
    for (long long n = (long long)blockIdx.x * RED_THREADS + threadIdx.x; n < n_nodes; n += (long long)gridDim.x * RED_THREADS) {
        const double r0 = r[3 * n], r1 = r[3 * n + 1], r2 = r[3 * n + 2];
        const double *d = dinv + 9 * n;
        const double z0 = d[0] * r0 + d[1] * r1 + d[2] * r2, z1 = d[3] * r0 + d[4] * r1 + d[5] * r2, z2 = d[6] * r0 + d[7] * r1 + d[8] * r2;
        z[3 * n] = z0; z[3 * n + 1] = z1; z[3 * n + 2] = z2;
        if (stage == ST_ZINIT) { p[3 * n] = z0; p[3 * n + 1] = z1; p[3 * n + 2] = z2; }
        acc += r0 * z0 + r1 * z1 + r2 * z2;
    }
    double v[1] = {acc}, t[1];
    if (grid_sum<1>(v, partials, 0, &s->ticket[0], t)) {
        sums[0] = t[0];
        if (fuse) { s->rz_new = t[0]; cg_scalar_step(s, stage); }
    }
}

// p = r + beta p  (skipped once converged: the reference returns before this update, cpu.jl:244-246)
__global__ void cg_p_kernel(long long n, double *__restrict__ p, const double *__restrict__ r, const CGScalars *s) {
    if (s->done) return;
    const double beta = s->beta;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = r[i] + beta * p[i];
}

__global__ void __launch_bounds__(RED_THREADS) dot_kernel(long long n, const double *__restrict__ a, const double *__restrict__ b,
                                                           double *partials, unsigned int *ticket, double *out) {
    double acc = 0;
    for (long long i = (long long)blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * RED_THREADS) acc += a[i] * b[i];
    double v[1] = {acc}, t[1];
    if (grid_sum<1>(v, partials, 2, ticket, t)) *out = t[0];
}

__global__ void axpby_kernel(long long n, double a, const double *__restrict__ x, double b, const double *__restrict__ y, double *__restrict__ z,
                             const uint8_t *__restrict__ fixed) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double v = a * (x ? x[i] : 0.0) + b * (y ? y[i] : 0.0);
        z[i] = (fixed && fixed[i]) ? 0.0 : v;
    }
}

static int ensure_cg_buffers(jfem_handle *h) {
    size_t n = (size_t)h->n_dofs();
    if (h->cg_r.n != n) {
        JFEM_TRY(h->cg_r.alloc(n)); JFEM_TRY(h->cg_p.alloc(n)); JFEM_TRY(h->cg_Ap.alloc(n));
    }
    if (h->cg_z.n != n) JFEM_TRY(h->cg_z.alloc(n));
    if (h->red_partials.n == 0) {
        JFEM_TRY(h->red_partials.alloc(4 * RED_BLOCKS + 8));
        JFEM_TRY(h->cg_s.alloc(1));
        JFEM_CUDA(cudaMemsetAsync(h->cg_s.p, 0, sizeof(CGScalars), h->stream));
    }
    return JFEM_OK;
}

static int allreduce_sums(jfem_handle *h, double *sums, int count) {
    if (h->n_ranks > 1) return comm_allreduce_sum(h, sums, count);
    return JFEM_OK;
}

int vec_dot(jfem_handle *h, const double *a, const double *b, double *out_host) {
    JFEM_TRY(ensure_cg_buffers(h));
    double *sums = h->red_partials.p + 4 * RED_BLOCKS;
    dot_kernel<<<RED_BLOCKS, RED_THREADS, 0, h->stream>>>(h->n_owned_dofs(), a, b, h->red_partials.p, &h->cg_s.p->ticket[2], sums + 4);
    JFEM_CUDA(cudaGetLastError());
    JFEM_TRY(allreduce_sums(h, sums + 4, 1));
    JFEM_CUDA(cudaMemcpyAsync(out_host, sums + 4, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    h->total_launches += 1;
    return JFEM_OK;
}

static int apply_operator(jfem_handle *h, int flags, double *x, double *y, const int *done) {
    if (h->n_ranks > 1) JFEM_TRY(halo_exchange(h, x, !(flags & JFEM_USE_CSR)));
    if (flags & JFEM_USE_CSR) return csr_spmv(h, x, y, flags | JFEM_PROJECT, done);
    return op_apply(h, (flags & JFEM_TANGENT) ? OP_TANGENT : OP_LINEAR, x, y, flags | JFEM_PROJECT, done);
}

int cg_solve(jfem_handle *h, const double *b, double *x, double tol, int rel, int max_iter, int flags, int *iters, double *resid) {
    JFEM_TRY(ensure_built(h));
    JFEM_TRY(ensure_cg_buffers(h));
    const long long n = h->n_dofs(), n_own = h->n_owned_dofs();
    const int fuse = h->n_ranks == 1 ? 1 : 0;
    double *sums = h->red_partials.p + 4 * RED_BLOCKS;
    CGScalars *s = h->cg_s.p;
    CGScalars init;
    memset(&init, 0, sizeof init);
    init.thr = tol; init.rel = rel; init.max_iter = max_iter;
    const int pcg = (flags & JFEM_JACOBI) ? 1 : 0;
    init.pcg = pcg;
    if (pcg) JFEM_TRY(jacobi_build(h, flags));
    const long long nn_own = n_own / 3;
    JFEM_CUDA(cudaMemcpyAsync(s, &init, sizeof init, cudaMemcpyHostToDevice, h->stream));
    // r0 = P(b - A x0)
    JFEM_TRY(apply_operator(h, flags, x, h->cg_Ap.p, nullptr));
    cg_init_kernel<<<RED_BLOCKS, RED_THREADS, 0, h->stream>>>(n_own, b, h->cg_Ap.p, h->fixed.p, h->cg_r.p, h->cg_p.p, h->red_partials.p, s, sums, fuse);
    JFEM_CUDA(cudaGetLastError());
    if (n > n_own) {   // ghost entries of r and p are refreshed by the halo exchange of p; keep them defined
        JFEM_CUDA(cudaMemsetAsync(h->cg_r.p + n_own, 0, (n - n_own) * sizeof(double), h->stream));
        JFEM_CUDA(cudaMemsetAsync(h->cg_p.p + n_own, 0, (n - n_own) * sizeof(double), h->stream));
    }
    if (!fuse) {
        JFEM_TRY(allreduce_sums(h, sums, 2));
        cg_scalar_kernel<<<1, 1, 0, h->stream>>>(s, ST_INIT, sums);
    }
    if (pcg) {
        pcg_z_kernel<<<RED_BLOCKS, RED_THREADS, 0, h->stream>>>(nn_own, h->cg_r.p, h->cg_dinv.p, h->cg_z.p, h->cg_p.p, h->red_partials.p, s, sums, fuse, ST_ZINIT);
        if (!fuse) { JFEM_TRY(allreduce_sums(h, sums, 1)); cg_scalar_kernel<<<1, 1, 0, h->stream>>>(s, ST_ZINIT, sums); }
    }
    h->total_launches += 2;
    int host_state[2] = {0, 0};   // iters, done
    JFEM_CUDA(cudaMemcpyAsync(host_state, &s->iters, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    const int chunk = 8;
    int launched = 0;
    while (launched < max_iter && !host_state[1]) {
        int todo = max_iter - launched < chunk ? max_iter - launched : chunk;
        for (int k = 0; k < todo; k++) {
            JFEM_TRY(apply_operator(h, flags, h->cg_p.p, h->cg_Ap.p, &s->done));
            cg_dot_kernel<<<RED_BLOCKS, RED_THREADS, 0, h->stream>>>(n_own, h->cg_p.p, h->cg_Ap.p, h->red_partials.p, s, sums, fuse);
            if (!fuse) { JFEM_TRY(allreduce_sums(h, sums, 1)); cg_scalar_kernel<<<1, 1, 0, h->stream>>>(s, ST_ALPHA, sums); }
            cg_update_kernel<<<RED_BLOCKS, RED_THREADS, 0, h->stream>>>(n_own, n_own, x, h->cg_r.p, h->cg_p.p, h->cg_Ap.p, h->red_partials.p, s, sums, fuse);
            if (!fuse) { JFEM_TRY(allreduce_sums(h, sums, 1)); cg_scalar_kernel<<<1, 1, 0, h->stream>>>(s, ST_UPDATE, sums); }
            if (pcg) {
                pcg_z_kernel<<<RED_BLOCKS, RED_THREADS, 0, h->stream>>>(nn_own, h->cg_r.p, h->cg_dinv.p, h->cg_z.p, h->cg_p.p, h->red_partials.p, s, sums, fuse, ST_Z);
                if (!fuse) { JFEM_TRY(allreduce_sums(h, sums, 1)); cg_scalar_kernel<<<1, 1, 0, h->stream>>>(s, ST_Z, sums); }
            }
            cg_p_kernel<<<RED_BLOCKS, RED_THREADS, 0, h->stream>>>(n_own, h->cg_p.p, pcg ? h->cg_z.p : h->cg_r.p, s);
            h->total_launches += (fuse ? 3 : 5) + (pcg ? (fuse ? 1 : 2) : 0);
        }
        JFEM_CUDA(cudaGetLastError());
        launched += todo;
        JFEM_CUDA(cudaMemcpyAsync(host_state, &s->iters, 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        JFEM_CUDA(cudaStreamSynchronize(h->stream));
        if (host_state[1]) break;
    }
    CGScalars fin;
    JFEM_CUDA(cudaMemcpyAsync(&fin, s, sizeof fin, cudaMemcpyDeviceToHost, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    if (iters) *iters = fin.iters;
    if (resid) *resid = sqrt(fin.rr);
    return JFEM_OK;
}

// Inexact Newton-Krylov (solve_newton_krylov_gpu!, ext/JuliaFEMCUDAExt.jl:685-854):
//   R = f_ext - f_int(u) (fixed dofs zeroed) ; stop if ||R|| < newton_tol ;
//   eta = min(forcing_max, ||R||^forcing_power), linear_tol = eta ||R|| (absolute) ; CG on K(u) du = R ; u += du.
// Deviations from the reference, both defects there: the operator is the true tangent K(u) (ext:468 subtracts
// f_ext inside the "operator"), and the Newton right-hand side is +R (ext:835 passes -R with R = f_ext - f_int).
// Plastic state is committed once per accepted Newton step (src/materials/abstract_material.jl:203-207).
int newton_krylov(jfem_handle *h, const double *fext, double *u, double newton_tol, int max_newton, int max_cg,
                  double forcing_power, double forcing_max, int flags, int *newton_iters, int *cg_iters, double *resid,
                  double *history, int history_cap) {
    JFEM_TRY(ensure_built(h));
    const long long n = h->n_dofs();
    if (h->nk_R.n != (size_t)n) { JFEM_TRY(h->nk_R.alloc(n)); JFEM_TRY(h->nk_du.alloc(n)); JFEM_TRY(h->nk_f.alloc(n)); }
    const int blocks = RED_BLOCKS, thr = 256;
    int total_cg = 0, it = 0;
    double Rn = 0;
    for (it = 0; it <= max_newton; it++) {
        if (h->n_ranks > 1) JFEM_TRY(halo_exchange(h, u, true));
        JFEM_TRY(op_apply(h, OP_RESIDUAL, u, h->nk_f.p, 0, nullptr));
        axpby_kernel<<<blocks, thr, 0, h->stream>>>(n, 1.0, fext, -1.0, h->nk_f.p, h->nk_R.p, h->fixed.p);
        JFEM_CUDA(cudaGetLastError());
        double rr;
        JFEM_TRY(vec_dot(h, h->nk_R.p, h->nk_R.p, &rr));
        int fail = 0;
        JFEM_CUDA(cudaMemcpy(&fail, h->dflags.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (fail) {
            JFEM_CUDA(cudaMemset(h->dflags.p, 0, sizeof(int)));
            jfem_set_error("Jacobian J = sqrt(det(C)) must be positive (invalid deformation during Newton step %d)", it);
            return JFEM_EDOMAIN;
        }
        Rn = sqrt(rr);
        if (Rn < newton_tol || it == max_newton) break;
        double eta = fmin(forcing_max, pow(Rn, forcing_power));
        double ltol = eta * Rn;
        JFEM_CUDA(cudaMemcpyAsync(h->ulin.p, u, n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        h->has_lin = true;
        if (flags & JFEM_USE_CSR) JFEM_TRY(csr_assemble(h, u, 1));
        JFEM_CUDA(cudaMemsetAsync(h->nk_du.p, 0, n * sizeof(double), h->stream));
        int ci = 0; double cr = 0;
        JFEM_TRY(cg_solve(h, h->nk_R.p, h->nk_du.p, ltol, 0, max_cg, flags | JFEM_TANGENT, &ci, &cr));
        total_cg += ci;
        if (history && it < history_cap) { history[3 * it] = ci; history[3 * it + 1] = Rn; history[3 * it + 2] = eta; }
        axpby_kernel<<<blocks, thr, 0, h->stream>>>(n, 1.0, u, 1.0, h->nk_du.p, u, nullptr);
        JFEM_CUDA(cudaGetLastError());
        h->total_launches += 2;
    }
    // accepted: commit the trial state computed by the last residual evaluation
    if (h->st_old.n && Rn < newton_tol) JFEM_CUDA(cudaMemcpyAsync(h->st_old.p, h->st_new.p, h->st_old.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    if (newton_iters) *newton_iters = it;
    if (cg_iters) *cg_iters = total_cg;
    if (resid) *resid = Rn;
    return JFEM_OK;
}

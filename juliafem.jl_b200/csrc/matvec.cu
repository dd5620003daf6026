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
    const uint8_t *blob;       // n_patches x stride bytes
    int n_patches, stride, off_pn, off_xl, off_xs, off_go, off_gs, off_lc;
    int max_nodes, max_nx;
    const double *coords;
    const double *x;
    const double *ulin;
    double *y;
    double *ipart;
    long long elem_offset;
    int project, atomic_iface, group_smem;
    int *fail;
    const int *done;
    long long *timing;         // debug: per-phase clock64 stamps of block 0 (nullptr = off)
};

// ---- mbarrier / TMA bulk-copy primitives (sm_90+; SASS: SYNCS.*, UBLKCP)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Persistent kernel: block b processes patches b, b+grid, b+2*grid, ...   Software pipeline per patch:
//   (a) wait for x / coordinates of the patch nodes: gathered asynchronously (cp.async) into shared memory during the
//       previous patch's phase 2, when the xs/Xs tiles are already free
//   (b) barrier; one thread issues the TMA bulk copy of the metadata blob of the patch after this one
//   (c) phase 1: one thread per element, register-resident contraction (elem.cuh); the 3*nnpe results are written
//       to a node-major staging tile (all contributions to one node contiguous, one plane per component)
//   (d) barrier; wait for the next blob; issue the asynchronous gather of the NEXT patch's x / coordinates
//   (e) phase 2: one thread per node: sum the node's contiguous run in fixed order (deterministic, no atomics);
//       interior nodes -> y, interface nodes -> one partial slot per (patch, node) for iface_reduce_kernel.
// Gathers and stores use the flat index i = 3*node + component so that a warp touches consecutive addresses wherever
// consecutive patch nodes have consecutive ids (the patch node lists are id-sorted).

// Ping-pong: a block holds G = 2 (T = 256) or 3 (T = 128) independent groups of T threads (when registers allow), each working on its own
// patch stream with its own shared-memory tiles.  Named barriers force the two groups to take turns in the fp64-bound
// phase 1, so that one group's gather / reduction / store phases always overlap the other group's arithmetic
// (two free-running blocks per SM were observed to run in lock-step instead: both in phase 1, then both in phase 2).
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Asynchronous gather (LDGSTS) of the x / coordinate / linearisation-point entries of the patch whose blob is at bl,
// straight into the shared tiles; no registers are held while the loads are in flight.  Flat index i = 3*node + component.
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int T, int NF>
__device__ __forceinline__ void patch_gather_async(const PatchKArgs &a, const unsigned char *bl, int tid, double *xs, double *Xs, double *us) {
    const int *hdr = reinterpret_cast<const int *>(bl);
    const int np = hdr[0], nx = (int)((unsigned)hdr[1] >> 16);
    const uint32_t *pn = reinterpret_cast<const uint32_t *>(bl + a.off_pn);
    const uint32_t *xl = reinterpret_cast<const uint32_t *>(bl + a.off_xl);
    for (int j = tid; j < np; j += T) {   // one node (3 x 8-byte copies) per thread: few instructions per copy
        const long long g = 3 * (long long)(pn[j] & PN_ID_MASK);
        cp_async8(xs + 3 * j, a.x + g); cp_async8(xs + 3 * j + 1, a.x + g + 1); cp_async8(xs + 3 * j + 2, a.x + g + 2);
        if (NF == 2) { cp_async8(us + 3 * j, a.ulin + g); cp_async8(us + 3 * j + 1, a.ulin + g + 1); cp_async8(us + 3 * j + 2, a.ulin + g + 2); }
    }
    for (int j = tid; j < nx; j += T) {
        const long long g = 3 * (long long)xl[j];
        cp_async8(Xs + 3 * j, a.coords + g); cp_async8(Xs + 3 * j + 1, a.coords + g + 1); cp_async8(Xs + 3 * j + 2, a.coords + g + 2);
    }
    cp_async_commit();
}

template <int NNPE, int CLS, int MODE>
struct PatchCfg {
    static constexpr bool fast = (CLS == CLASS_AFFINE && MODE == OP_LINEAR && NNPE == 10) || NNPE == 4;
};

template <int NNPE, int CLS, int MODE, class Pt, int T, int G>
__global__ void __launch_bounds__(T * G, 1) patch_kernel(PatchKArgs a, Pt pt) {
    extern __shared__ __align__(128) unsigned char smraw_all[];
    if (a.done && *a.done) return;
    constexpr int NF = Pt::NF;
    constexpr int PS = NNPE * T + 5;   // plane stride of the staging tile (odd offset: the 3 planes fall in different banks)
    const int grp = threadIdx.x / T, tid = threadIdx.x - grp * T;
    unsigned char *smraw = smraw_all + (size_t)grp * a.group_smem;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smraw);
    unsigned char *blob0 = smraw + 16;
    double *stage = reinterpret_cast<double *>(smraw + 16 + 2 * (size_t)a.stride);
    double *xs = stage + 3 * PS + 1;
    double *Xs = xs + 3 * a.max_nodes;
    double *us = Xs + 3 * a.max_nx;
    const int stride_p = gridDim.x * G;   // patches are dealt round-robin over blocks first (balanced per SM), then groups

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (G > 1 && grp == G - 1) named_arrive(8, 2 * T);   // group 0 takes the first turn in phase 1
    int p = blockIdx.x + grp * gridDim.x;
    if (tid == 0 && p < a.n_patches) {
        mbar_expect_tx(&mbar[0], a.stride);
        bulk_g2s(blob0, a.blob + (size_t)p * a.stride, a.stride, &mbar[0]);
    }
    if (p < a.n_patches) {
        mbar_wait(&mbar[0], 0);
        patch_gather_async<T, NF>(a, blob0, tid, xs, Xs, us);
    }
    for (int it = 0; p < a.n_patches; it++, p += stride_p) {
        const int buf = it & 1;
        const unsigned char *bl = blob0 + (size_t)buf * a.stride;
        const int *hdr = reinterpret_cast<const int *>(bl);
        const int np = hdr[0], nif = hdr[1] & 0xFFFF, nx = (int)((unsigned)hdr[1] >> 16), ipb = hdr[2], ne = hdr[3];
        const uint32_t *pn = reinterpret_cast<const uint32_t *>(bl + a.off_pn);
        const uint16_t *xsl = reinterpret_cast<const uint16_t *>(bl + a.off_xs);
        if (a.timing && blockIdx.x == 0 && threadIdx.x == 0 && it < 8) a.timing[it * 8 + 0] = clock64();

        // ---- (a) the asynchronous gather of this patch (issued during the previous patch's phase 2) must have landed
        cp_async_wait_all();
        if (a.timing && blockIdx.x == 0 && threadIdx.x == 0 && it < 8) a.timing[it * 8 + 1] = clock64();
        if (G == 1) __syncthreads(); else named_sync(1 + grp, T);
        if (a.timing && blockIdx.x == 0 && threadIdx.x == 0 && it < 8) a.timing[it * 8 + 2] = clock64();
        // ---- (b) every thread has left phase 2 of the previous patch: its blob buffer may be refilled
        const bool has_next = p + stride_p < a.n_patches;
        if (tid == 0 && has_next) {
            mbar_expect_tx(&mbar[buf ^ 1], a.stride);
            bulk_g2s(blob0 + (size_t)(buf ^ 1) * a.stride, a.blob + (size_t)(p + stride_p) * a.stride, a.stride, &mbar[buf ^ 1]);
        }

        // ---- (c) phase 1: element contraction (the two groups alternate here)
        if (G > 1) named_sync(8 + grp, 2 * T);
        const uint16_t *go = reinterpret_cast<const uint16_t *>(bl + a.off_go);
        if (tid < ne) {
            const uint16_t *lc = reinterpret_cast<const uint16_t *>(bl + a.off_lc);
            const uint8_t *rk = reinterpret_cast<const uint8_t *>(bl + a.off_gs);
            int n[NNPE];
            JF_UNROLL for (int k = 0; k < NNPE; k++) n[k] = lc[k * T + tid];
            auto out = [=](int k, double v0, double v1, double v2) {
                double *d = stage + go[n[k]] + rk[k * T + tid];
                d[0] = v0; d[PS] = v1; d[2 * PS] = v2;
            };
            bool ok = true;
            const long long el = a.elem_offset + (long long)p * T + tid;
            if (CLS == CLASS_AFFINE && MODE == OP_LINEAR && NNPE == 10) {
                int nxs[4];
                JF_UNROLL for (int k = 0; k < 4; k++) nxs[k] = xsl[n[k]];
                SField U{xs, n};
                SField X{Xs, nxs};
                PtLinear q; q.la = pt.la; q.mu = pt.mu; q.sy = 0; q.H = 0; q.pe = pt.pe; q.pe_n = pt.pe_n;
                q.load(el);
                tet10_affine_linear(q.la, q.mu, U, X, out);
            } else {
                int nxs[NNPE];
                JF_UNROLL for (int k = 0; k < NNPE; k++) nxs[k] = xsl[n[k]];
                SField X{Xs, nxs};
                SField F[NF];
                F[0].base = xs; F[0].n = n;
                if (NF == 2) { F[NF - 1].base = us; F[NF - 1].n = n; }
                if (NNPE == 10) ok = tet10_general(pt, el, F, X, out);
                else if (NNPE == 8) ok = hex8_general(pt, el, F, X, out);
                else ok = tet4_general(pt, el, F, X, out);
            }
            if (!ok) atomicOr(a.fail, 1);
        }
        if (a.timing && blockIdx.x == 0 && threadIdx.x == 0 && it < 8) a.timing[it * 8 + 3] = clock64();
        if (G > 1) named_arrive(8 + (grp + 1) % G, 2 * T);
        if (G == 1) __syncthreads(); else named_sync(1 + grp, T);
        if (a.timing && blockIdx.x == 0 && threadIdx.x == 0 && it < 8) a.timing[it * 8 + 4] = clock64();

        // ---- (d) start the next patch's gather (its blob was requested at (b))
        if (has_next) {
            mbar_wait(&mbar[buf ^ 1], ((it + 1) >> 1) & 1);
            if (a.timing && blockIdx.x == 0 && threadIdx.x == 0 && it < 8) a.timing[it * 8 + 7] = clock64();
            patch_gather_async<T, NF>(a, blob0 + (size_t)(buf ^ 1) * a.stride, tid, xs, Xs, us);
        }
        if (a.timing && blockIdx.x == 0 && threadIdx.x == 0 && it < 8) a.timing[it * 8 + 5] = clock64();

        // ---- (e) phase 2: one thread per node: ordered reduction of the node's contiguous staged run (3 planes)
        for (int j = tid; j < np; j += T) {
            const int q0 = go[j], q1 = go[j + 1];
            const double *sp = stage + q0;
            const int cnt = q1 - q0;
            double s0 = 0, s1 = 0, s2 = 0;
            int q = 0;
            for (; q + 2 <= cnt; q += 2) {
                const double a0 = sp[q], a1 = sp[q + 1], b0 = sp[PS + q], b1 = sp[PS + q + 1], c0 = sp[2 * PS + q], c1 = sp[2 * PS + q + 1];
                s0 += a0; s1 += b0; s2 += c0;
                s0 += a1; s1 += b1; s2 += c1;
            }
            if (q < cnt) { s0 += sp[q]; s1 += sp[PS + q]; s2 += sp[2 * PS + q]; }
            const uint32_t w = pn[j];
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
                    double *pp = a.ipart + 3 * ((long long)ipb + j);
                    pp[0] = s0; pp[1] = s1; pp[2] = s2;
                }
            } else {
                double *py = a.y + 3 * (long long)(w & PN_ID_MASK);
                py[0] = s0; py[1] = s1; py[2] = s2;
            }
        }
        if (a.timing && blockIdx.x == 0 && threadIdx.x == 0 && it < 8) a.timing[it * 8 + 6] = clock64();
        // no barrier here: the in-flight gather writes only xs/Xs/us, which phase 2 does not read
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Warp-specialised variant for the headline case (affine Tet10, linear elastic): one persistent block per SM with
//   * 8 COMPUTE warps (256 threads, registers raised with setmaxnreg): phase 1 of patch i, back to back
//   * 12 HELPER warps (384 threads, registers lowered): asynchronous gather of patch i+1 and the per-node reduction +
//     stores of patch i-1, concurrently with the compute warps
// so the fp64 pipe never waits for the LSU-bound phases.  Hand-off through mbarriers (full/empty pairs on the double
// buffered x tile and staging tile); metadata blobs triple-buffered by TMA bulk copies.
// ------------------------------------------------------------------------------------------------------------------
#define WS_T 256
#define WS_H 384
#define WS_NB 4             // metadata blob ring (TMA), refilled two patches ahead of first use
#define WS_COMPUTE_REGS 168 // 256*168 + 384*48 <= 640*96: setmaxnreg redistributes the launch allocation, it cannot grow it
#define WS_HELPER_REGS 48
#define WS_GX 9             // flat x items per compute thread held in registers by the look-ahead gather (768 nodes)
#define WS_GC 2             // flat coordinate items per compute thread

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct WsSmem {   // byte offsets inside dynamic shared memory, computed on the host
    int blob, stage, xs, Xs, total;
};

// COMPUTE warps, per patch i:  issue the global loads of patch i+1's x / coordinates into registers (coalesced LDG: the
// patch node lists are id-sorted) -> phase 1 of patch i out of the x tile (i & 1) into staging tile (i & 1) -> signal the
// helpers -> park the prefetched registers in x tile ((i+1) & 1).  The load latency hides behind phase 1.
// HELPER warps, per patch k: wait for staging tile (k & 1) -> ordered per-node reduction + stores -> release the tile ->
// refill blob slot (k % WS_NB) with patch k + WS_NB by one TMA bulk copy.
template <class Pt>
__global__ void __launch_bounds__(WS_T + WS_H, 1) patch_kernel_ws(PatchKArgs a, Pt pt, WsSmem L) {
    extern __shared__ __align__(128) unsigned char sm[];
    if (a.done && *a.done) return;
    constexpr int NNPE = 10, T = WS_T, PS = NNPE * T + 5;
    uint64_t *mb = reinterpret_cast<uint64_t *>(sm);
    uint64_t *blob_full = mb, *stage_full = mb + WS_NB, *stage_empty = mb + WS_NB + 2;
    unsigned char *blobs = sm + L.blob;
    const size_t stage_sz = 3 * PS + 1;
    double *stage0 = reinterpret_cast<double *>(sm + L.stage);
    double *xs0 = reinterpret_cast<double *>(sm + L.xs);
    double *Xs0 = reinterpret_cast<double *>(sm + L.Xs);
    const int xs_sz = 3 * a.max_nodes, Xs_sz = 3 * a.max_nx;
    const int n_it = (a.n_patches - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int k = 0; k < WS_NB; k++) mbar_init(&blob_full[k], 1);
        for (int k = 0; k < 2; k++) { mbar_init(&stage_full[k], WS_T); mbar_init(&stage_empty[k], WS_H); }
        mbar_fence_init();
        for (int k = 0; k < WS_NB && k < n_it; k++) {
            mbar_expect_tx(&blob_full[k], a.stride);
            bulk_g2s(blobs + (size_t)k * a.stride, a.blob + (size_t)(blockIdx.x + k * gridDim.x) * a.stride, a.stride, &blob_full[k]);
        }
    }
    __syncthreads();

    if (threadIdx.x < WS_T) {
        // =========================== compute warps ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_COMPUTE_REGS));
        const int tid = threadIdx.x;
        // look-ahead gather registers: flat items i = 3*node + component, thread t holds items t + r*T (consecutive
        // lanes -> consecutive addresses wherever consecutive patch nodes have consecutive ids)
        double rx[WS_GX], rc[WS_GC];
        JF_UNROLL for (int r = 0; r < WS_GX; r++) rx[r] = 0.0;
        JF_UNROLL for (int r = 0; r < WS_GC; r++) rc[r] = 0.0;
        auto load_regs = [&](const unsigned char *bl) {
            const int *hdr = reinterpret_cast<const int *>(bl);
            const int n3 = 3 * hdr[0], nx3 = 3 * (int)((unsigned)hdr[1] >> 16);
            const uint32_t *pn = reinterpret_cast<const uint32_t *>(bl + a.off_pn);
            const uint32_t *xl = reinterpret_cast<const uint32_t *>(bl + a.off_xl);
            JF_UNROLL for (int r = 0; r < WS_GX; r++) {
                const int i = tid + r * T, ic = i < n3 ? i : 0, j = ic / 3;
                rx[r] = __ldg(a.x + 3 * (long long)(pn[j] & PN_ID_MASK) + (ic - 3 * j));
            }
            JF_UNROLL for (int r = 0; r < WS_GC; r++) {
                const int i = tid + r * T, ic = i < nx3 ? i : 0, j = ic / 3;
                rc[r] = __ldg(a.coords + 3 * (long long)(nx3 > 0 ? xl[j] : 0) + (ic - 3 * j));
            }
        };
        // registers -> shared tiles (+ slow paths for patches with more than WS_GX*T/3 nodes)
        auto store_regs = [&](const unsigned char *bl, double *xs, double *Xs) {
            const int *hdr = reinterpret_cast<const int *>(bl);
            const int n3 = 3 * hdr[0], nx3 = 3 * (int)((unsigned)hdr[1] >> 16);
            JF_UNROLL for (int r = 0; r < WS_GX; r++) {
                const int i = tid + r * T;
                if (i < n3) xs[i] = rx[r];
            }
            JF_UNROLL for (int r = 0; r < WS_GC; r++) {
                const int i = tid + r * T;
                if (i < nx3) Xs[i] = rc[r];
            }
            const uint32_t *pn = reinterpret_cast<const uint32_t *>(bl + a.off_pn);
            const uint32_t *xl = reinterpret_cast<const uint32_t *>(bl + a.off_xl);
            for (int i = tid + WS_GX * T; i < n3; i += T) {
                const int j = i / 3;
                xs[i] = __ldg(a.x + 3 * (long long)(pn[j] & PN_ID_MASK) + (i - 3 * j));
            }
            for (int i = tid + WS_GC * T; i < nx3; i += T) {
                const int j = i / 3;
                Xs[i] = __ldg(a.coords + 3 * (long long)xl[j] + (i - 3 * j));
            }
        };
        if (n_it > 0) {
            mbar_wait(&blob_full[0], 0);
            load_regs(blobs);
            store_regs(blobs, xs0, Xs0);
            named_sync(2, WS_T);
        }
        for (int i = 0; i < n_it; i++) {
            const unsigned char *bl = blobs + (size_t)(i % WS_NB) * a.stride;
            const unsigned char *bln = blobs + (size_t)((i + 1) % WS_NB) * a.stride;
            double *stage = stage0 + (size_t)(i & 1) * stage_sz;
            const double *xs = xs0 + (size_t)(i & 1) * xs_sz, *Xs = Xs0 + (size_t)(i & 1) * Xs_sz;
            const bool has_next = i + 1 < n_it;
            const bool tm = a.timing && blockIdx.x == 0 && tid == 0 && i < 8;
            if (tm) a.timing[i * 8 + 0] = clock64();
            if (has_next) {
                mbar_wait(&blob_full[(i + 1) % WS_NB], ((i + 1) / WS_NB) & 1);
                if (tm) a.timing[i * 8 + 5] = clock64();
                load_regs(bln);
            }
            if (tm) a.timing[i * 8 + 1] = clock64();
            if (i >= 2) mbar_wait(&stage_empty[i & 1], ((i >> 1) - 1) & 1);
            if (tm) a.timing[i * 8 + 2] = clock64();
            const int ne = reinterpret_cast<const int *>(bl)[3];
            if (tid < ne) {
                const uint16_t *go = reinterpret_cast<const uint16_t *>(bl + a.off_go);
                const uint16_t *xsl = reinterpret_cast<const uint16_t *>(bl + a.off_xs);
                const uint16_t *lc = reinterpret_cast<const uint16_t *>(bl + a.off_lc);
                const uint8_t *rk = reinterpret_cast<const uint8_t *>(bl + a.off_gs);
                int n[NNPE], nxs[4];
                JF_UNROLL for (int k = 0; k < NNPE; k++) n[k] = lc[k * T + tid];
                JF_UNROLL for (int k = 0; k < 4; k++) nxs[k] = xsl[n[k]];
                auto out = [=](int k, double v0, double v1, double v2) {
                    double *d = stage + go[n[k]] + rk[k * T + tid];
                    d[0] = v0; d[PS] = v1; d[2 * PS] = v2;
                };
                SField U{xs, n};
                SField X{Xs, nxs};
                Pt q = pt;
                q.load(a.elem_offset + (long long)(blockIdx.x + i * gridDim.x) * T + tid);
                tet10_affine_linear(q.la, q.mu, U, X, out);
            }
            if (tm) a.timing[i * 8 + 3] = clock64();
            mbar_arrive(&stage_full[i & 1]);
            if (has_next) {
                store_regs(bln, xs0 + (size_t)((i + 1) & 1) * xs_sz, Xs0 + (size_t)((i + 1) & 1) * Xs_sz);
                named_sync(2, WS_T);   // tile (i+1)&1 complete and visible to every compute warp
            }
            if (tm) a.timing[i * 8 + 4] = clock64();
        }
    } else {
        // =========================== helper warps ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_HELPER_REGS));
        const int hid = threadIdx.x - WS_T;
        for (int k = 0; k < n_it; k++) {
            const unsigned char *bl = blobs + (size_t)(k % WS_NB) * a.stride;
            const double *stage = stage0 + (size_t)(k & 1) * stage_sz;
            const bool tm = a.timing && blockIdx.x == 0 && hid == 0 && k < 8;
            if (tm) a.timing[64 + k * 8 + 0] = clock64();
            mbar_wait(&blob_full[k % WS_NB], (k / WS_NB) & 1);
            mbar_wait(&stage_full[k & 1], (k >> 1) & 1);
            if (tm) a.timing[64 + k * 8 + 1] = clock64();
            const int *hdr = reinterpret_cast<const int *>(bl);
            const int np = hdr[0], nif = hdr[1] & 0xFFFF, ipb = hdr[2];
            const uint32_t *pn = reinterpret_cast<const uint32_t *>(bl + a.off_pn);
            const uint16_t *go = reinterpret_cast<const uint16_t *>(bl + a.off_go);
            for (int j = hid; j < np; j += WS_H) {
                const int q0 = go[j], cnt = go[j + 1] - q0;
                const double *sp = stage + q0;
                double s0 = 0, s1 = 0, s2 = 0;
                int q = 0;
                for (; q + 2 <= cnt; q += 2) {
                    const double a0 = sp[q], a1 = sp[q + 1], b0 = sp[PS + q], b1 = sp[PS + q + 1], c0 = sp[2 * PS + q], c1 = sp[2 * PS + q + 1];
                    s0 += a0; s1 += b0; s2 += c0;
                    s0 += a1; s1 += b1; s2 += c1;
                }
                if (q < cnt) { s0 += sp[q]; s1 += sp[PS + q]; s2 += sp[2 * PS + q]; }
                const uint32_t w = pn[j];
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
                        double *pp = a.ipart + 3 * ((long long)ipb + j);
                        pp[0] = s0; pp[1] = s1; pp[2] = s2;
                    }
                } else {
                    double *py = a.y + 3 * (long long)(w & PN_ID_MASK);
                    py[0] = s0; py[1] = s1; py[2] = s2;
                }
            }
            if (tm) a.timing[64 + k * 8 + 2] = clock64();
            mbar_arrive(&stage_empty[k & 1]);
            // blob slot (k % WS_NB) is free once every helper thread is past the reduction: refill it with patch k + WS_NB
            if (k + WS_NB < n_it) {
                named_sync(1, WS_H);
                if (hid == 0) {
                    mbar_expect_tx(&blob_full[k % WS_NB], a.stride);
                    bulk_g2s(blobs + (size_t)(k % WS_NB) * a.stride, a.blob + (size_t)(blockIdx.x + (k + WS_NB) * gridDim.x) * a.stride, a.stride,
                             &blob_full[k % WS_NB]);
                }
            }
        }
    }
}

// y[interface node] = sum of its partial slots, ascending (set, patch) order.  3 threads per node; the first four
// slots come from one 16-byte load (islot4), rarer nodes with more slots continue through the CSR list.
__global__ void iface_reduce_kernel(const uint32_t *__restrict__ inodes, const int4 *__restrict__ islot4, const int32_t *__restrict__ iptr,
                                    const int32_t *__restrict__ islots, const double *__restrict__ ipart,
                                    double *__restrict__ y, long long n3, const int *done) {
    if (done && *done) return;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    const int node = (int)(i / 3), c = (int)(i - 3LL * node);
    const uint32_t w = inodes[node];
    const int4 s4 = islot4[node];
    double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
    if (s4.x >= 0) v0 = __ldcg(ipart + 3LL * s4.x + c);
    if (s4.y >= 0) v1 = __ldcg(ipart + 3LL * s4.y + c);
    if (s4.z >= 0) v2 = __ldcg(ipart + 3LL * s4.z + c);
    if (s4.w >= 0) v3 = __ldcg(ipart + 3LL * s4.w + c);
    double s = 0;
    if (s4.x >= 0) s += v0;
    if (s4.y >= 0) s += v1;
    if (s4.z >= 0) s += v2;
    if (s4.w >= 0) s += v3;
    if (w & PN_NEEDX)   // overflow flag: more than four patches touch this node
        for (int q = iptr[node] + 4; q < iptr[node + 1]; q++) s += __ldcg(ipart + 3LL * islots[q] + c);
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
        PatchSetHost &S = h->hsets[c];
        if (S.n_elems == 0) continue;
        std::vector<uint32_t> w = S.pnodes;
        apply_fixed_words(h, w);
        for (int p = 0; p < S.n_patches; p++)
            memcpy(&S.blob[(size_t)p * S.stride + S.off_pn], &w[S.pnode_ptr[p]], 4 * (size_t)(S.pnode_ptr[p + 1] - S.pnode_ptr[p]));
        JFEM_TRY(h->dsets[c].blob.upload(S.blob));
    }
    std::vector<uint32_t> w = h->hif.inodes;
    apply_fixed_words(h, w);
    for (size_t i = 0; i < w.size(); i++)
        if (h->hif.iptr[i + 1] - h->hif.iptr[i] > 4) w[i] |= PN_NEEDX; else w[i] &= ~PN_NEEDX;   // bit 27 = overflow flag here
    JFEM_TRY(h->inodes.upload(w));
    JFEM_TRY(h->fixed.upload(h->mesh.fixed));
    return JFEM_OK;
}

// per-element parameters: caller order (n_params x n_elems) -> SoA [4][n_elems] in internal element order
int upload_material(jfem_handle *h) {
    if (h->mat_per_elem.empty()) { h->matp.release(); return JFEM_OK; }
    const int64_t ne = h->mesh.n_elems;
    std::vector<double> soa(4 * (size_t)ne, 0.0);
    int64_t off = 0;
    for (int c = 0; c < N_CLASSES; c++) {
        const PatchSetHost &S = h->hsets[c];
        for (int64_t i = 0; i < S.n_elems; i++)
            for (int q = 0; q < h->mat_nparams; q++) soa[(size_t)q * ne + off + i] = h->mat_per_elem[(size_t)S.elem_perm[i] * h->mat_nparams + q];
        off += S.n_elems;
    }
    return h->matp.upload(soa);
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
        D.max_nx = S.max_nx; D.off_pn = S.off_pn; D.off_xl = S.off_xl; D.off_xs = S.off_xs; D.off_go = S.off_go; D.off_gs = S.off_gs; D.off_lc = S.off_lc;
        D.stride = S.stride;
    }
    {
        const size_t ni = h->hif.inodes.size();
        std::vector<int32_t> s4(4 * ni, -1);
        for (size_t i = 0; i < ni; i++)
            for (int q = h->hif.iptr[i], k = 0; q < h->hif.iptr[i + 1] && k < 4; q++, k++) s4[4 * i + k] = h->hif.islots[q];
        JFEM_TRY(h->islot4.upload(s4));
    }
    JFEM_TRY(h->e2i.upload(e2i));
    JFEM_TRY(h->iptr.upload(h->hif.iptr));
    JFEM_TRY(h->islots.upload(h->hif.islots));
    JFEM_TRY(h->ipart.alloc((size_t)3 * h->hif.n_partials + 3));
    JFEM_TRY(h->coords.upload(h->mesh.coords));
    JFEM_TRY(h->dflags.alloc(4));
    JFEM_CUDA(cudaMemset(h->dflags.p, 0, 4 * sizeof(int)));
    h->built = true;
    JFEM_TRY(upload_fixed(h));
    JFEM_TRY(upload_material(h));
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
    if constexpr (NNPE == 10 && CLS == CLASS_AFFINE && MODE == OP_LINEAR && T == WS_T) {
        if (h->warp_specialised) {
            constexpr int PS = NNPE * T + 5;
            auto r128 = [](size_t v) { return (v + 127) & ~(size_t)127; };
            WsSmem L;
            L.blob = 128;
            L.stage = (int)r128(L.blob + WS_NB * (size_t)D.stride);
            L.xs = (int)r128(L.stage + 2 * sizeof(double) * (3 * PS + 1));
            L.Xs = (int)r128(L.xs + 2 * sizeof(double) * 3 * D.max_nodes);
            L.total = (int)r128(L.Xs + 2 * sizeof(double) * 3 * D.max_nx);
            if (L.total <= 227 * 1024) {
                auto kws = patch_kernel_ws<Pt>;
                static int configured_ws = 0;
                if (L.total > configured_ws) {
                    JFEM_CUDA(cudaFuncSetAttribute(kws, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
                    configured_ws = L.total;
                }
                h->last_smem = L.total; h->last_blocks_per_sm = 1;
                int grid = h->n_sms < D.n_patches ? h->n_sms : D.n_patches;
                kws<<<grid, WS_T + WS_H, L.total, h->stream>>>(a, pt, L);
                JFEM_CUDA(cudaGetLastError());
                h->matvec_launches++;
                return JFEM_OK;
            }
        }
    }
    constexpr int G = !PatchCfg<NNPE, CLS, MODE>::fast ? 1 : (T == 128 ? 3 : (T == 256 ? 2 : 1));
    size_t gsm = 16 + 2 * (size_t)D.stride + sizeof(double) * (3 * (NNPE * T + 5) + 1 + (size_t)3 * D.max_nodes * (NF == 2 ? 2 : 1) + (size_t)3 * D.max_nx);
    gsm = (gsm + 127) & ~(size_t)127;
    a.group_smem = (int)gsm;
    size_t smem = gsm * G;
    if (G > 1 && smem > 227 * 1024) { jfem_set_error("patch needs %zu bytes of shared memory; lower patch_elems", smem); return JFEM_EINVAL; }
    auto kern = patch_kernel<NNPE, CLS, MODE, Pt, T, G>;
    static size_t configured = 0;
    if (smem > configured) {
        JFEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    if (smem > 227 * 1024) { jfem_set_error("patch needs %zu bytes of shared memory; lower patch_elems", smem); return JFEM_EINVAL; }
    int per_sm = 1;
    JFEM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T * G, smem));
    if (per_sm < 1) per_sm = 1;
    h->last_smem = (int64_t)smem; h->last_blocks_per_sm = per_sm;
    int grid = h->n_sms * per_sm;
    if (grid * G > D.n_patches) grid = (D.n_patches + G - 1) / G;
    kern<<<grid, T * G, smem, h->stream>>>(a, pt);
    JFEM_CUDA(cudaGetLastError());
    h->matvec_launches++;
    return JFEM_OK;
}

void fill_material(const jfem_handle *h, jf::MatBase &m) {
    m.la = h->mat[0] * h->mat[1] / ((1.0 + h->mat[1]) * (1.0 - 2.0 * h->mat[1]));   // linear_elastic.jl:82
    m.mu = h->mat[0] / (2.0 * (1.0 + h->mat[1]));                                    // :97
    m.sy = h->mat[2]; m.H = h->mat[3];
    m.pe = h->matp.n ? h->matp.p : nullptr;
    m.pe_n = h->mesh.n_elems;
}

template <int NNPE, int CLS, int T>
static int dispatch_mode(jfem_handle *h, const PatchSetDev &D, PatchKArgs a, int mode) {
    const long long n_gp = (long long)h->mesh.n_elems * h->ngp();
    if (h->mat_kind == JFEM_MAT_LINEAR_ELASTIC || (mode == OP_LINEAR)) {
        // small-strain linear elasticity: K.x = f_int(x), tangent == K
        PtLinear pt; fill_material(h, pt);
        return launch_set<NNPE, CLS, OP_LINEAR, PtLinear, T>(h, D, a, pt);
    }
    if (h->mat_kind == JFEM_MAT_NEO_HOOKEAN) {
        if (mode == OP_RESIDUAL) { PtNHResidual pt; fill_material(h, pt); return launch_set<NNPE, CLASS_GENERAL, OP_RESIDUAL, PtNHResidual, T>(h, D, a, pt); }
        PtNHTangent pt; fill_material(h, pt);
        return launch_set<NNPE, CLASS_GENERAL, OP_TANGENT, PtNHTangent, T>(h, D, a, pt);
    }
    if (h->mat_kind == JFEM_MAT_PERFECT_PLASTICITY) {
        if (mode == OP_RESIDUAL) {
            PtPPResidual pt; fill_material(h, pt);
            pt.st_old = h->st_old.p; pt.st_new = h->st_new.p; pt.n_gp = n_gp;
            return launch_set<NNPE, CLASS_GENERAL, OP_RESIDUAL, PtPPResidual, T>(h, D, a, pt);
        }
        PtPPTangent pt; fill_material(h, pt);
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
        a.blob = D.blob.p; a.n_patches = D.n_patches; a.stride = D.stride; a.off_pn = D.off_pn; a.off_xl = D.off_xl; a.off_xs = D.off_xs; a.off_go = D.off_go;
        a.off_gs = D.off_gs; a.off_lc = D.off_lc; a.max_nx = D.max_nx;
        a.coords = h->coords.p; a.x = x; a.ulin = h->ulin.p; a.y = y; a.ipart = h->ipart.p;
        a.elem_offset = D.elem_offset; a.max_nodes = D.max_nodes;
        a.project = (flags & JFEM_PROJECT) ? 1 : 0; a.atomic_iface = atomic_iface;
        a.fail = h->dflags.p; a.done = done; a.timing = h->timing.p;
        int rc;
        switch (h->mesh.nnpe) {
            case 10: rc = dispatch_threads<10>(h, D, a, mode); break;
            case 8: rc = dispatch_threads<8>(h, D, a, mode); break;
            default: rc = dispatch_threads<4>(h, D, a, mode); break;
        }
        JFEM_TRY(rc);
    }
    if (!atomic_iface && n3) {
        iface_reduce_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, h->stream>>>(h->inodes.p, (const int4 *)h->islot4.p, h->iptr.p, h->islots.p, h->ipart.p, y, n3, done);
        JFEM_CUDA(cudaGetLastError());
        h->matvec_launches++;
    }
    h->total_launches += h->matvec_launches;
    return JFEM_OK;
}

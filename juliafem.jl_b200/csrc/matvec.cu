// matvec.cu -- the matrix-free element operator  y = A(x)  on patches (one persistent thread block per SM streams patches).
//
// Replaces compute_element_stresses_kernel! + nodal_assembly_kernel! (ext/JuliaFEMCUDAExt.jl:222-361) and the
// atomic Tet10 residual kernel of demos/gpu_assembly_tet10.jl:145-232.  Per patch:
//   gather   x (and coordinates, linearisation point) of the patch's nodes into a shared tile (coalesced, id-sorted)
//   phase 1  one thread per element: register-resident element contraction (elem.cuh); the 3*nnpe results are scattered
//            into the jagged staging tile through precomputed entry indices (patches.cpp: lanes are assigned so that a
//            half-warp hits distinct banks as often as possible)
//   phase 2  three lanes per patch node (one per component): sum the node's staged contributions row by row in a fixed
//            order (deterministic, no atomics; consecutive lanes read consecutive words).  Interior nodes are stored to
//            y; an interface node's partial goes to its slot of this patch (the slots of a node are contiguous) and
//            iface_reduce_kernel adds them in ascending patch order.  (A fused "last patch to arrive sums" variant
//            was measured 3x SLOWER: the fence -> ticket atomic -> fence -> load chain sits in every helper warp.)
// Metadata arrives as one blob per patch by TMA bulk copies (layout: common.h).
#include <chrono>
#include <cstring>
#include <type_traits>

#include "handle.h"
#include "patch_elem.cuh"

using namespace jf;

struct PatchKArgs {
    const uint8_t *blob;       // n_patches x stride bytes
    int n_patches, stride;
    int offB, offC, bytesA, bytesB, bytesC;   // the three parts of a blob (A starts at 0)
    int a_cx, a_cX, b_et, c_qn, c_ql, c_jo;   // table offsets relative to their part
    int max_nodes, max_entries, x_all;
    int xs_cap, Xs_cap;        // doubles per x tile / coordinate tile (2 per gather chunk)
    long long n_dofs;          // length of x / coordinates in doubles (a chunk that would read past it is copied per double)
    const double *coords;
    const double *x;
    const double *ulin;
    double *y;
    double *ipart;
    const uint32_t *slot_node; // partial slot -> node id (only read by the non-deterministic atomic mode)
    long long elem_offset;
    int project, atomic_iface, nbuf;
    int dbg;                   // debug_skip bits: 1 = no global stores in phase 2, 2 = no row loops, 4 = no phase 1, 8 = no look-ahead gather
    int *fail;
    const int *done;
    long long *timing;         // debug: per-phase clock64 stamps of block 0 (nullptr = off)
    // in-kernel interface reduction (cooperative launch: every block is resident): after a grid-wide barrier all
    // threads add the partial slots of the interface nodes.  tail = 0 -> iface_reduce_kernel does it instead.
    int tail;
    unsigned int *gbar;        // monotonic arrival counter
    unsigned int gbar_target;  // value of the counter once every block of THIS launch has arrived
    const uint32_t *inodes;
    const int32_t *ibase;
    long long n_inodes3;
    // halo exchange fused into the ws kernel (partitioned meshes): the helper warps push this rank's interface values into
    // the neighbours' landing buffers during the pipeline prologue; patches that read ghost values are ordered last and
    // wait for the neighbours' flags; ghost values are read straight from the landing buffer.
    int halo;
    uint32_t n_owned;          // nodes >= n_owned are ghosts (0xFFFFFFFF: none)
    HaloFused hf;
};

// y[interface node] = sum of its partial slots (contiguous, ascending (set, patch) order); one thread per node, three
// components (one index lookup per node, 24 contiguous bytes per slot).  Nodes no element touches are listed with an empty
// slot range (y = 0).
__device__ __forceinline__ void iface_item(const uint32_t *__restrict__ inodes, const int32_t *__restrict__ ibase, const double *__restrict__ ipart,
                                           double *__restrict__ y, long long node) {
    const int b0 = ibase[node], b1 = ibase[node + 1];
    const double *pp = ipart + 3LL * b0;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    int r = 0;
    for (; r + 2 <= b1 - b0; r += 2) {
        const double a0 = __ldcg(pp + 3 * r), a1 = __ldcg(pp + 3 * r + 1), a2 = __ldcg(pp + 3 * r + 2);
        const double c0 = __ldcg(pp + 3 * r + 3), c1 = __ldcg(pp + 3 * r + 4), c2 = __ldcg(pp + 3 * r + 5);
        s0 += a0; s1 += a1; s2 += a2;
        s0 += c0; s1 += c1; s2 += c2;
    }
    if (r < b1 - b0) { s0 += __ldcg(pp + 3 * r); s1 += __ldcg(pp + 3 * r + 1); s2 += __ldcg(pp + 3 * r + 2); }
    double *d = y + 3LL * inodes[node];
    d[0] = s0; d[1] = s1; d[2] = s2;
}

// Grid-wide barrier + interface reduction at the end of a (cooperatively launched) patch kernel.  Called by all threads.
__device__ __forceinline__ void iface_tail(const PatchKArgs &a) {
    __syncthreads();   // every thread of the block has issued its partial stores
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.gbar, 1u);
        while ((int)(*(volatile unsigned int *)a.gbar - a.gbar_target) < 0) __nanosleep(40);
        __threadfence();
    }
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_inodes3 / 3; i += (long long)gridDim.x * blockDim.x)
        iface_item(a.inodes, a.ibase, a.ipart, a.y, i);
}

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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Asynchronous gather (LDGSTS.128) of the 16-byte chunks a patch reads from x (and ulin / the coordinates) straight into the
// shared tiles; no registers are held while the copies are in flight.  Part A of the blob lists the chunks in ascending
// order, so the tile is the concatenation of the chunks and the element table addresses it by offset.
__device__ __forceinline__ void cp_async8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// the executing thread's prior cp.async operations arrive on the mbarrier when they have landed (count pre-charged: .noinc)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Gather the chunks j = tid, tid + nthr, ... of a chunk list into `tile` (chunk c = doubles 2c, 2c+1): src = dof vector (or
// coordinates) of n_tot doubles, of which the first n_own are read from src itself by one 128-bit asynchronous copy per
// chunk; a chunk that ends beyond n_own is copied per double: doubles n_own..n_tot-1 (ghost values of a partitioned mesh,
// fused halo) come from the landing buffer `land` the neighbours wrote (volatile: the same addresses are rewritten every
// second exchange, no stale L1 line).  Returns true if this thread wrote the tile with ordinary stores.
__device__ __forceinline__ bool gather_chunks(const uint32_t *ch, int n, int tid, int nthr, double *tile, const double *src, long long n_own,
                                              long long n_tot, const double *land) {
    bool plain = false;
    for (int j = tid; j < n; j += nthr) {
        const long long d = 2LL * ch[j];
        if (d + 2 <= n_own) cp_async16(tile + 2 * j, src + d);
        else {
            JF_UNROLL for (int e = 0; e < 2; e++) {
                if (d + e < n_own) cp_async8(tile + 2 * j + e, src + d + e);
                else if (d + e < n_tot && land) { tile[2 * j + e] = __ldcv(land + (d + e - n_own)); plain = true; }
            }
        }
    }
    return plain;
}

// All T threads of a block: issue the gather of the patch whose part A is at pa and arrive on `bar` (count T).
template <int T, int NF>
__device__ __forceinline__ void patch_gather_async(const PatchKArgs &a, const unsigned char *pa, int tid, double *xs, double *Xs, double *us, uint64_t *bar) {
    const int *hdr = reinterpret_cast<const int *>(pa);
    const int ncx = hdr[1] & 0xFFFF, ncX = hdr[1] >> 16;
    const uint32_t *cx = reinterpret_cast<const uint32_t *>(pa + a.a_cx);
    const uint32_t *cX = reinterpret_cast<const uint32_t *>(pa + a.a_cX);
    gather_chunks(cx, ncx, tid, T, xs, a.x, a.n_dofs, a.n_dofs, nullptr);
    if (NF == 2) gather_chunks(cx, ncx, tid, T, us, a.ulin, a.n_dofs, a.n_dofs, nullptr);
    if (a.x_all) gather_chunks(cx, ncx, tid, T, Xs, a.coords, a.n_dofs, a.n_dofs, nullptr);
    else gather_chunks(cX, ncX, tid, T, Xs, a.coords, a.n_dofs, a.n_dofs, nullptr);
    cp_async_mbar_arrive(bar);
}

// ---- phase 2 of one patch by NT threads (t = 0..NT-1): ten nodes per warp, lane = 3 * node + component.
// pc = part C of the blob, stage = the staging tile phase 1 filled.
template <int NT>
__device__ __forceinline__ void reduce_patch(const PatchKArgs &a, const unsigned char *pc, const double *stage, int t) {
    const int np = reinterpret_cast<const int *>(pc)[0];
    const uint32_t *qn = reinterpret_cast<const uint32_t *>(pc + a.c_qn);
    const uint8_t *ql = pc + a.c_ql;
    const uint16_t *jo = reinterpret_cast<const uint16_t *>(pc + a.c_jo);
    const int warp = t >> 5, lane = t & 31;
    const int sub = lane / 3, c = lane - 3 * sub;
    for (int base = warp * 10; base < np; base += (NT / 32) * 10) {
        const int q = base + sub;
        const bool valid = sub < 10 && q < np;
        double s = 0.0;
        uint32_t w = 0;
        if (valid) {
            w = qn[q];
            const int len = ql[q];
            const double *sp = stage + 3 * q + c;
            int r = 0;
            for (; r + 4 <= len; r += 4) {
                const double v0 = sp[3 * jo[r]], v1 = sp[3 * jo[r + 1]], v2 = sp[3 * jo[r + 2]], v3 = sp[3 * jo[r + 3]];
                s += v0; s += v1; s += v2; s += v3;
            }
            for (; r < len; r++) s += sp[3 * jo[r]];
            if (a.project && ((w >> (PN_FIXSHIFT + c)) & 1u)) s = 0.0;
        }
        const bool ifc = valid && (w & PN_IFACE);
        const long long e = 3 * (long long)(w & PN_ID_MASK) + c;
        if (valid && !ifc) a.y[e] = s;
        if (ifc) {
            if (a.atomic_iface) atomicAdd(a.y + 3 * (long long)a.slot_node[w & PN_ID_MASK] + c, s);
            else a.ipart[e] = s;   // summed with the node's other slots by iface_reduce_kernel
        }
    }
}

// ---- phase 2 for the warp-specialised kernel, "flat" form.  The staging tile is read as rows of 3 * rowlen[r] consecutive
// doubles (component-interleaved, nodes in reduce order q); flat index i = 3 q + c.  Thread t of the NT helper threads owns
// the flat indices t, t + NT, t + 2 NT, ... (MC accumulators, "chains") and adds row after row: consecutive lanes read
// consecutive words (conflict-free, all 32 lanes busy).  The row lengths are non-increasing (nodes sorted by descending
// contribution count), so the rows in which a warp has more than MC/2, more than MC/4 or at least one chain are three
// contiguous blocks: their bounds come from one ballot each over row descriptors held one per lane, and each block is a
// counted loop over predicated loads with no data-dependent branch inside (the descriptors travel by shuffle, not through
// memory), unrolled so that several rows are in flight.  Summation order per (node, component): rows ascending -- fixed,
// no atomics, bitwise reproducible.
// rows [r, r_end) of the current block of 32 row descriptors, C predicated chains, U rows in flight
template <int NT, int C, int U>
__device__ __forceinline__ void flat_rows(const double *sp, unsigned desc, int &r, int r_end, int rb, int base, int t, double *acc) {
#pragma unroll U
    for (; r < r_end; r++) {
        const unsigned d = __shfl_sync(0xFFFFFFFFu, desc, r - rb);
        const double *row = sp + 3 * (int)(d & 0xFFFFu);
        const int l3 = (int)(d >> 16) - base;
        JF_UNROLL for (int m = 0; m < C; m++)
            if (t + NT * m < l3) acc[m] += row[NT * m];
    }
}

template <int NT, int MC>
__device__ __forceinline__ void reduce_patch_flat(const PatchKArgs &a, const unsigned char *pc, const double *stage, int t, long long *tm = nullptr) {
    const int *hdr = reinterpret_cast<const int *>(pc);
    const int np = hdr[0], nrows = (a.dbg & 2) ? 0 : (hdr[3] & 0xFFFF), n3 = 3 * np;
    const uint32_t *qn = reinterpret_cast<const uint32_t *>(pc + a.c_qn);
    const uint16_t *jo = reinterpret_cast<const uint16_t *>(pc + a.c_jo);
    const int lane = t & 31, wf = t & ~31;   // wf = first flat index of this warp inside a block of NT
    for (int base = 0; base < n3; base += NT * MC) {   // one pass unless the patch has more than NT * MC / 3 nodes
        double acc[MC];
        JF_UNROLL for (int m = 0; m < MC; m++) acc[m] = 0.0;
        const double *sp = stage + base + t;
        if (tm) tm[3] = clock64();
        for (int rb = 0; rb < nrows; rb += 32) {         // row descriptors, one per lane: offset | flat length << 16
            unsigned desc = 0;
            if (rb + lane < nrows) {
                const unsigned o = jo[rb + lane];
                desc = o | ((3u * ((unsigned)jo[rb + lane + 1] - o)) << 16);
            }
            const int l3 = (int)(desc >> 16) - base - wf;   // flat entries of my row from this warp's first index on
            // row blocks by the number of chains this warp has in them (non-increasing): > MC/2, > MC/4, >= 1
            const int R8 = rb + __popc(__ballot_sync(0xFFFFFFFFu, l3 > NT * (MC / 2)));
            const int R4 = rb + __popc(__ballot_sync(0xFFFFFFFFu, l3 > NT * (MC / 4)));
            const int R2 = rb + __popc(__ballot_sync(0xFFFFFFFFu, l3 > 0));
            int r = rb;
            flat_rows<NT, MC, 2>(sp, desc, r, R8, rb, base, t, acc);
            flat_rows<NT, MC / 2, 2>(sp, desc, r, R4, rb, base, t, acc);
            flat_rows<NT, MC / 4, 4>(sp, desc, r, R2, rb, base, t, acc);
        }
        if (tm) { tm[4] = clock64(); tm[6] = nrows; }
        if (a.dbg & 1) continue;
        // stores: all table lookups first, then one predicated store per chain
        uint32_t w[MC];
        JF_UNROLL for (int m = 0; m < MC; m++) {
            const int i = base + t + NT * m;
            w[m] = qn[i < n3 ? (int)(((unsigned)i * 43691u) >> 17) : 0];   // i / 3 for i < 98 304
        }
        JF_UNROLL for (int m = 0; m < MC; m++) {
            const int i = base + t + NT * m;
            const int c = i - 3 * (int)(((unsigned)i * 43691u) >> 17);
            double v = acc[m];
            if (a.project && ((w[m] >> (PN_FIXSHIFT + c)) & 1u)) v = 0.0;
            const long long id = w[m] & PN_ID_MASK;
            if (i < n3) {
                if (!(w[m] & PN_IFACE)) a.y[3 * id + c] = v;
                else if (!a.atomic_iface) a.ipart[3 * id + c] = v;   // interface node: its partial slot of this patch
                else atomicAdd(a.y + 3 * (long long)a.slot_node[id] + c, v);
            }
        }
    }
}

// Generic persistent kernel (every element type / material / operator mode): block b processes patches b, b+grid, ...
// Software pipeline per patch:
//   (a) wait for x / coordinates of the patch nodes: gathered asynchronously (cp.async) during the previous patch's phase 2
//   (b) barrier; one thread issues the TMA bulk copy of the metadata blob of the patch after this one
//   (c) phase 1   (d) barrier; wait for the next blob; issue the asynchronous gather of the NEXT patch   (e) phase 2
template <int NNPE, int CLS, int MODE, class Pt, int T>
__global__ void __launch_bounds__(T, 1) patch_kernel(PatchKArgs a, Pt pt) {
    extern __shared__ __align__(128) unsigned char smraw[];
    if (a.done && *a.done) {   // converged (device-side flag): nothing to do, but keep the barrier counter in step with the host
        if (a.tail && threadIdx.x == 0) atomicAdd(a.gbar, 1u);
        return;
    }
    constexpr int NF = Pt::NF;
    const int tid = threadIdx.x;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smraw);
    unsigned char *blob0 = smraw + 32;
    double *stage = reinterpret_cast<double *>(smraw + 32 + a.nbuf * (size_t)a.stride);
    double *xs = stage + 3 * a.max_entries;
    double *Xs = xs + a.xs_cap;
    double *us = Xs + a.Xs_cap;
    const int stride_p = gridDim.x;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_init(&mbar[2], T);   // x / coordinate tiles of the next patch have landed
        mbar_fence_init();
    }
    __syncthreads();
    int p = blockIdx.x;
    if (tid == 0 && p < a.n_patches) {
        mbar_expect_tx(&mbar[0], a.stride);
        bulk_g2s(blob0, a.blob + (size_t)p * a.stride, a.stride, &mbar[0]);
    }
    if (p < a.n_patches) {
        mbar_wait(&mbar[0], 0);
        patch_gather_async<T, NF>(a, blob0, tid, xs, Xs, us, &mbar[2]);
    }
    for (int it = 0; p < a.n_patches; it++, p += stride_p) {
        const int buf = a.nbuf == 2 ? (it & 1) : 0;
        const unsigned char *bl = blob0 + (size_t)buf * a.stride;
        const int ne = reinterpret_cast<const int *>(bl)[2];
        // ---- (a) the asynchronous gather of this patch (issued during the previous patch's phase 2) must have landed
        mbar_wait(&mbar[2], it & 1);
        __syncthreads();
        // ---- (b) every thread has left phase 2 of the previous patch: its blob buffer may be refilled
        const bool has_next = p + stride_p < a.n_patches;
        if (tid == 0 && has_next && a.nbuf == 2) {
            mbar_expect_tx(&mbar[buf ^ 1], a.stride);
            bulk_g2s(blob0 + (size_t)(buf ^ 1) * a.stride, a.blob + (size_t)(p + stride_p) * a.stride, a.stride, &mbar[buf ^ 1]);
        }
        // ---- (c) phase 1: element contraction
        if (tid < ne) {
            const uint32_t *et = reinterpret_cast<const uint32_t *>(bl + a.offB + a.b_et);
            const bool ok = element_phase<NNPE, CLS, MODE, Pt, T>(pt, a.elem_offset + (long long)p * T + tid, et, tid, xs, Xs, us, stage);
            if (!ok) atomicOr(a.fail, 1);
        }
        __syncthreads();
        // ---- (d) start the next patch's gather (its blob was requested at (b))
        if (has_next && a.nbuf == 2) {
            mbar_wait(&mbar[buf ^ 1], ((it + 1) >> 1) & 1);
            patch_gather_async<T, NF>(a, blob0 + (size_t)(buf ^ 1) * a.stride, tid, xs, Xs, us, &mbar[2]);
        }
        // ---- (e) phase 2 (the in-flight gather writes only xs/Xs/us, which phase 2 does not read)
        reduce_patch<T>(a, bl + a.offC, stage, tid);
        // ---- (f) single blob buffer (very large patches): fetch the next blob only now
        if (has_next && a.nbuf == 1) {
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(&mbar[0], a.stride);
                bulk_g2s(blob0, a.blob + (size_t)(p + stride_p) * a.stride, a.stride, &mbar[0]);
            }
            mbar_wait(&mbar[0], (it + 1) & 1);
            patch_gather_async<T, NF>(a, blob0, tid, xs, Xs, us, &mbar[2]);
        }
    }
    if (a.tail) iface_tail(a);
}

// ------------------------------------------------------------------------------------------------------------------
// Warp-specialised variant (linear-elastic operators on affine elements): one persistent block per SM with two roles
//   * 8 COMPUTE warps (256 threads, registers raised with setmaxnreg): phase 1 of patch i and nothing else.  No barrier
//     between the compute warps: they wait on mbarriers only and drift apart freely (two warps in lockstep on one
//     scheduler stall together on the shared fp64 pipe: 64 % vs 75 % pipe utilisation in isolation,
//     profiles/microbench/phase1_occupancy.cu)
//   * 8 HELPER warps (256 threads, registers lowered): for patch k, first the look-ahead gather of patch k+2 -- 128-bit
//     asynchronous copies (LDGSTS.128) of the 16-byte chunks of x and of the coordinate array listed in part A of the
//     blob, straight into the x tile patch k has just released; completion is signalled on an mbarrier by
//     cp.async.mbarrier.arrive, nobody waits on the loads -- then phase 2 of patch k, concurrently with phase 1 of patch
//     k+1.  They also request the blob parts (TMA bulk copies) and, on partitioned meshes, carry the halo exchange.
// Hand-off through mbarrier full/empty pairs on double-buffered x tiles and staging tiles; the three blob parts are
// double buffered separately (they are live at different times).
// ------------------------------------------------------------------------------------------------------------------
#define WS_T 256            // compute threads = elements per patch
#define WS_H 256            // helper threads
#define WS_COMPUTE_REGS 168 // 256*168 + 256*88 = 512*128: setmaxnreg redistributes the launch allocation, it cannot grow it
#define WS_HELPER_REGS 88
#define WS_MC 8             // flat indices per helper thread and pass (256 x 8 = 2 048 >= 3 x 682 nodes)

struct WsSmem {   // byte offsets inside dynamic shared memory, computed on the host
    int A, B, C, stage, xs, Xs, total;
    int strideA, strideB, strideC;
};

template <int NNPE, int CLS, int MODE, class Pt>
__global__ void __launch_bounds__(WS_T + WS_H, 1) patch_kernel_ws(PatchKArgs a, Pt pt, WsSmem L) {
    extern __shared__ __align__(128) unsigned char sm[];
    constexpr int T = WS_T;
    uint64_t *mb = reinterpret_cast<uint64_t *>(sm);
    uint64_t *A_full = mb, *B_full = mb + 2, *C_full = mb + 4, *stage_full = mb + 6, *stage_empty = mb + 8, *x_full = mb + 10;
    const size_t stage_sz = 3 * (size_t)a.max_entries;
    double *stage0 = reinterpret_cast<double *>(sm + L.stage);
    double *xs0 = reinterpret_cast<double *>(sm + L.xs);
    double *Xs0 = reinterpret_cast<double *>(sm + L.Xs);
    const int n_it = (a.n_patches - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto patch_of = [&](int i) { return (size_t)(blockIdx.x + (size_t)i * gridDim.x); };
    auto reqA = [&](int i) {
        mbar_expect_tx(&A_full[i & 1], a.bytesA);
        bulk_g2s(sm + L.A + (size_t)(i & 1) * L.strideA, a.blob + patch_of(i) * a.stride, a.bytesA, &A_full[i & 1]);
    };
    auto reqB = [&](int i) {
        mbar_expect_tx(&B_full[i & 1], a.bytesB);
        bulk_g2s(sm + L.B + (size_t)(i & 1) * L.strideB, a.blob + patch_of(i) * a.stride + a.offB, a.bytesB, &B_full[i & 1]);
    };
    auto reqC = [&](int i) {
        mbar_expect_tx(&C_full[i & 1], a.bytesC);
        bulk_g2s(sm + L.C + (size_t)(i & 1) * L.strideC, a.blob + patch_of(i) * a.stride + a.offC, a.bytesC, &C_full[i & 1]);
    };

    if (threadIdx.x == 0) {
        for (int k = 0; k < 6; k++) mbar_init(&mb[k], 1);
        for (int k = 0; k < 2; k++) { mbar_init(&stage_full[k], WS_T); mbar_init(&stage_empty[k], WS_H); mbar_init(&x_full[k], WS_H); }
        mbar_fence_init();
        for (int k = 0; k < 2 && k < n_it; k++) { reqA(k); reqB(k); reqC(k); }
    }
    __syncthreads();
    if (a.done && *a.done) {   // device-side convergence flag of the CG loop: nothing to do; let the requested copies land first
        for (int k = 0; k < 2 && k < n_it; k++) { mbar_wait(&A_full[k], 0); mbar_wait(&B_full[k], 0); mbar_wait(&C_full[k], 0); }
        if (a.tail && threadIdx.x == 0) atomicAdd(a.gbar, 1u);
        return;
    }

    if (threadIdx.x < WS_T) {
        // =========================== compute warps ===========================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WS_COMPUTE_REGS));
        const int tid = threadIdx.x;
        for (int i = 0; i < n_it; i++) {
            const int b = i & 1;
            const bool tm = a.timing && blockIdx.x == 0 && tid == 0 && i < 8;
            if (tm) a.timing[i * 8 + 0] = clock64();
            mbar_wait(&x_full[b], (i >> 1) & 1);
            mbar_wait(&B_full[b], (i >> 1) & 1);
            if (tm) a.timing[i * 8 + 1] = clock64();
            if (i >= 2) mbar_wait(&stage_empty[b], ((i >> 1) - 1) & 1);
            if (tm) a.timing[i * 8 + 2] = clock64();
            const unsigned char *pb = sm + L.B + (size_t)b * L.strideB;
            const int ne = reinterpret_cast<const int *>(pb)[2];
            if (tid < ne && !(a.dbg & 4)) {
                const uint32_t *et = reinterpret_cast<const uint32_t *>(pb + a.b_et);
                const bool ok = element_phase<NNPE, CLS, MODE, Pt, T>(pt, a.elem_offset + (long long)patch_of(i) * T + tid, et, tid,
                                                                      xs0 + (size_t)b * a.xs_cap, Xs0 + (size_t)b * a.Xs_cap, nullptr,
                                                                      stage0 + (size_t)b * stage_sz);
                if (!ok) atomicOr(a.fail, 1);
            }
            if (tm) a.timing[i * 8 + 3] = clock64();
            mbar_arrive(&stage_full[b]);   // release: this thread's staging stores; x tile b and part B slot b are free, too
        }
    } else {
        // =========================== helper warps ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WS_HELPER_REGS));
        const int hid = threadIdx.x - WS_T;
        if (a.halo) {
            // push this rank's interface values into the neighbours' landing buffers (all helper threads of all blocks),
            // fence, and let the last block to finish publish the sequence flags -- while the pipeline fills
            const long long n3 = 3 * a.hf.send_off[a.hf.n_nb];
            for (long long i = (long long)blockIdx.x * WS_H + hid; i < n3; i += (long long)gridDim.x * WS_H) {
                const long long q = i / 3;
                int nb = 0;
                while (q >= a.hf.send_off[nb + 1]) nb++;
                a.hf.peer_land[nb][i - 3 * a.hf.send_off[nb]] = a.x[3LL * a.hf.send_nodes[q] + (i - 3 * q)];
            }
            // one system-scope fence per block, by the thread that takes the ticket (the barrier orders the other threads'
            // remote stores before it; fence cumulativity makes them visible before the flag).
            named_sync(1, WS_H);
            if (hid == 0) {
                __threadfence_system();
                const unsigned int t = atomicAdd(a.hf.ticket, 1u);
                if (t == gridDim.x - 1) {
                    *a.hf.ticket = 0;
                    __threadfence_system();
                    for (int nb = 0; nb < a.hf.n_nb; nb++) *((volatile unsigned long long *)a.hf.peer_flag[nb]) = a.hf.seq;
                }
            }
        }
        bool halo_seen = !a.halo;
        const long long n_own3 = a.halo ? 3LL * a.n_owned : a.n_dofs;   // doubles of x that are this rank's own
        // look-ahead gather of patch j into x tile j & 1 (part A slot j & 1 holds its chunk lists)
        auto gather = [&](int j) {
            const int b = j & 1;
            mbar_wait(&A_full[b], (j >> 1) & 1);
            const unsigned char *pa = sm + L.A + (size_t)b * L.strideA;
            const int *hdr = reinterpret_cast<const int *>(pa);
            const int ncx = (a.dbg & 8) ? 0 : (hdr[1] & 0xFFFF), ncX = (a.dbg & 8) ? 0 : (hdr[1] >> 16);
            if (!halo_seen && (hdr[3] & 0x10000)) {   // first patch that reads ghost values: the neighbours' data must have landed
                if (hid < a.hf.n_nb) {
                    const volatile unsigned long long *f = a.hf.my_flag[hid];
                    while (*f < a.hf.seq) { }
                    __threadfence_system();   // acquire side: only the polling threads fence; the barrier extends it to the block
                }
                named_sync(1, WS_H);
                halo_seen = true;
            }
            const uint32_t *cx = reinterpret_cast<const uint32_t *>(pa + a.a_cx);
            const uint32_t *cX = reinterpret_cast<const uint32_t *>(pa + a.a_cX);
            double *xs = xs0 + (size_t)b * a.xs_cap, *Xs = Xs0 + (size_t)b * a.Xs_cap;
            const bool plain = gather_chunks(cx, ncx, hid, WS_H, xs, a.x, n_own3, a.n_dofs, a.halo ? a.hf.land : nullptr);
            if (a.x_all) gather_chunks(cx, ncx, hid, WS_H, Xs, a.coords, a.n_dofs, a.n_dofs, nullptr);
            else gather_chunks(cX, ncX, hid, WS_H, Xs, a.coords, a.n_dofs, a.n_dofs, nullptr);
            if (plain) __threadfence_block();
            cp_async_mbar_arrive(&x_full[b]);
        };
        for (int j = 0; j < 2 && j < n_it; j++) gather(j);
        if (n_it > 2) {
            named_sync(1, WS_H);   // every helper thread has read the chunk lists of patches 0 and 1
            if (hid == 0) { reqA(2); if (n_it > 3) reqA(3); }
        }
        for (int k = 0; k < n_it; k++) {
            const bool tm = a.timing && blockIdx.x == 0 && hid == 0 && k < 8;
            if (tm) a.timing[64 + k * 8 + 0] = clock64();
            mbar_wait(&C_full[k & 1], (k >> 1) & 1);
            mbar_wait(&stage_full[k & 1], (k >> 1) & 1);
            if (tm) a.timing[64 + k * 8 + 1] = clock64();
            // every compute thread is past phase 1 of patch k: x tile and part B slot k & 1 are free -> patch k + 2
            if (k + 2 < n_it) {
                if (hid == 0) reqB(k + 2);
                gather(k + 2);
            }
            if (tm) a.timing[64 + k * 8 + 5] = clock64();
            reduce_patch_flat<WS_H, WS_MC>(a, sm + L.C + (size_t)(k & 1) * L.strideC, stage0 + (size_t)(k & 1) * stage_sz, hid, tm ? a.timing + 64 + k * 8 : nullptr);
            if (tm) a.timing[64 + k * 8 + 2] = clock64();
            mbar_arrive(&stage_empty[k & 1]);
            // part C slot (k & 1) is free once every helper thread is past the reduction, part A slot (k & 1) once every
            // helper thread has issued the gather of patch k + 2: refill them with patches k + 2 / k + 4
            if (k + 2 < n_it) {
                named_sync(1, WS_H);
                if (hid == 0) { reqC(k + 2); if (k + 4 < n_it) reqA(k + 4); }
            }
        }
        cp_async_wait_all();   // nothing may still be in flight into shared memory when the block exits
    }
    if (a.tail) iface_tail(a);
}

// y[interface node] = sum of its partial slots (contiguous, ascending (set, patch) order).  One thread per node.
// Nodes no element touches are listed with an empty slot range (y = 0).
__global__ void iface_reduce_kernel(const uint32_t *__restrict__ inodes, const int32_t *__restrict__ ibase, const double *__restrict__ ipart,
                                    double *__restrict__ y, long long n3, const int *done) {
    // launched with programmatic stream serialisation: the grid may be set up while the patch kernel drains; everything
    // the patch kernel wrote is visible after this call
    cudaGridDependencySynchronize();
    if (done && *done) return;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (3 * i < n3) iface_item(inodes, ibase, ipart, y, i);
}

__global__ void nodes_zero_kernel(const uint32_t *__restrict__ nodes, double *__restrict__ y, long long n3, const int *done) {
    if (done && *done) return;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    const int node = (int)(i / 3), c = (int)(i - 3LL * node);
    y[3LL * (nodes[node] & PN_ID_MASK) + c] = 0.0;
}

// ------------------------------------------------------------------------------------------------ build

int upload_fixed(jfem_handle *h) {   // (re)upload everything that embeds the Dirichlet mask
    if (!h->built) return JFEM_OK;
    for (int c = 0; c < N_CLASSES; c++) {
        PatchSetHost &S = h->hsets[c];
        if (S.n_elems == 0) continue;
        for (int p = 0; p < S.n_patches; p++) {
            uint32_t *qn = reinterpret_cast<uint32_t *>(&S.blob[(size_t)p * S.L.stride + S.L.off_qn]);
            for (int q = 0, nb = S.pnode_ptr[p]; q < S.pnode_ptr[p + 1] - nb; q++) {
                uint32_t w = S.qnodes[nb + q] & ~(7u << PN_FIXSHIFT);
                const int64_t id = S.qids[nb + q];
                for (int d = 0; d < 3; d++)
                    if (h->mesh.fixed[3 * id + d]) w |= (1u << (PN_FIXSHIFT + d));
                qn[q] = w;
            }
        }
        JFEM_TRY(h->dsets[c].blob.upload(S.blob));
    }
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
    // closed-form (affine) element classes only for the linear-elastic operator: the general kernels index the coordinate
    // tile by x-tile position, which the affine patch sets do not provide
    const bool use_affine = h->affine && (h->mat_kind == JFEM_MAT_LINEAR_ELASTIC || h->mat_kind < 0);
    classify_elements(h->mesh, use_affine);
    JFEM_TRY(build_patch_sets(h->mesh, h->patch_elems, use_affine, h->lane_window, h->n_ranks > 1 ? h->n_owned_nodes : -1, h->hsets, h->hif));
    h->setup_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int64_t off = 0;
    std::vector<int64_t> e2i(h->mesh.n_elems);
    for (int c = 0; c < N_CLASSES; c++) {
        PatchSetHost &S = h->hsets[c];
        PatchSetDev &D = h->dsets[c];
        D.release();
        D.cls = c; D.nnpe = S.nnpe; D.EP = S.EP; D.nxr = S.nxr; D.n_patches = S.n_patches; D.max_nodes = S.max_nodes; D.max_ncx = S.max_ncx; D.max_ncX = S.max_ncX;
        D.max_rows = S.max_rows; D.max_entries = S.max_entries; D.L = S.L;
        D.n_elems = S.n_elems; D.elem_offset = off;
        for (int64_t i = 0; i < S.n_elems; i++) e2i[S.elem_perm[i]] = off + i;
        off += S.n_elems;
    }
    JFEM_TRY(h->e2i.upload(e2i));
    {   // interface nodes followed by the nodes no element touches (empty slot range -> y = 0)
        std::vector<uint32_t> nodes = h->hif.inodes;
        std::vector<int32_t> base = h->hif.ibase;
        base.push_back((int32_t)h->hif.n_partials);
        for (uint32_t n : h->hif.orphans) { nodes.push_back(n); base.push_back((int32_t)h->hif.n_partials); }
        JFEM_TRY(h->inodes.upload(nodes));
        JFEM_TRY(h->ibase.upload(base));
    }
    JFEM_TRY(h->ipart.alloc((size_t)3 * h->hif.n_partials + 3));
    {
        std::vector<uint32_t> sn((size_t)h->hif.n_partials + 1, 0);
        for (size_t i = 0; i < h->hif.inodes.size(); i++) {
            const int64_t e = i + 1 < h->hif.inodes.size() ? h->hif.ibase[i + 1] : h->hif.n_partials;
            for (int64_t q = h->hif.ibase[i]; q < e; q++) sn[q] = h->hif.inodes[i];
        }
        JFEM_TRY(h->slot_node.upload(sn));
    }
    JFEM_TRY(h->gbar.alloc(4));
    JFEM_CUDA(cudaMemset(h->gbar.p, 0, 4 * sizeof(unsigned int)));
    h->gbar_count = 0;
    {
        int coop = 0;
        JFEM_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device));
        h->coop_ok = coop != 0;
    }
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

// Opt a kernel into `bytes` of dynamic shared memory.  The attribute is per device, so the cache lives in the handle
// (one handle = one device), not in a function-local static.
static int ensure_dynamic_smem(jfem_handle *h, const void *func, size_t bytes) {
    size_t &have = h->smem_attr[func];
    if (bytes > have) {
        JFEM_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        have = bytes;
    }
    return JFEM_OK;
}

static bool ws_layout(const PatchSetDev &D, int x_all, WsSmem &L) {
    auto r128 = [](size_t v) { return (v + 127) & ~(size_t)127; };
    const size_t xtile = sizeof(double) * 2 * (size_t)D.max_ncx, Xtile = sizeof(double) * 2 * (size_t)(x_all ? D.max_ncx : D.max_ncX);
    L.strideA = (int)r128(D.L.bytesA()); L.strideB = (int)r128(D.L.bytesB()); L.strideC = (int)r128(D.L.bytesC());
    L.A = 128;
    L.B = L.A + 2 * L.strideA;
    L.C = L.B + 2 * L.strideB;
    L.stage = L.C + 2 * L.strideC;
    L.xs = (int)r128(L.stage + 2 * sizeof(double) * 3 * (size_t)D.max_entries);
    L.Xs = (int)r128(L.xs + 2 * xtile);
    L.total = (int)r128(L.Xs + 2 * Xtile);
    return L.total <= 227 * 1024;
}


template <int NNPE, int CLS, int MODE, class Pt, int T>
static int launch_set(jfem_handle *h, const PatchSetDev &D, PatchKArgs a, const Pt &pt) {
    constexpr int NF = Pt::NF;
    auto r128 = [](size_t v) { return (v + 127) & ~(size_t)127; };
    const size_t Xtile = sizeof(double) * (size_t)a.Xs_cap;
    if constexpr (MODE == OP_LINEAR && T == WS_T && NF == 1 && (NNPE == 10 || NNPE == 8) && CLS == CLASS_AFFINE) {
        if (h->warp_specialised) {
            WsSmem L;
            if (ws_layout(D, a.x_all, L)) {
                auto kws = patch_kernel_ws<NNPE, CLS, MODE, Pt>;
                JFEM_TRY(ensure_dynamic_smem(h, (const void *)kws, (size_t)L.total));
                h->last_smem = L.total; h->last_blocks_per_sm = 1;
                int grid = h->n_sms < D.n_patches ? h->n_sms : D.n_patches;
                if (a.tail) {
                    a.gbar_target = (h->gbar_count += (unsigned)grid);
                    void *args[] = {(void *)&a, (void *)&pt, (void *)&L};
                    JFEM_CUDA(cudaLaunchCooperativeKernel((const void *)kws, dim3(grid), dim3(WS_T + WS_H), args, (size_t)L.total, h->stream));
                } else {
                    // (launching this kernel itself with programmatic stream serialisation was measured: -4 % per CG
                    // iteration on T1 but +10 % on T10, where the early-resident persistent blocks starve the tail of
                    // the preceding vector kernel -- not used)
                    kws<<<grid, WS_T + WS_H, L.total, h->stream>>>(a, pt, L);
                }
                JFEM_CUDA(cudaGetLastError());
                h->matvec_launches++;
                return JFEM_OK;
            }
        }
    }
    const size_t tiles = sizeof(double) * (3 * (size_t)D.max_entries + (size_t)a.xs_cap * (NF == 2 ? 2 : 1)) + Xtile;
    a.nbuf = 2;
    size_t smem = r128(32 + 2 * (size_t)D.L.stride + tiles);
    if (smem > 227 * 1024) { a.nbuf = 1; smem = r128(32 + (size_t)D.L.stride + tiles); }
    if (smem > 227 * 1024) { jfem_set_error("patch needs %zu bytes of shared memory; lower patch_elems", smem); return JFEM_EINVAL; }
    auto kern = patch_kernel<NNPE, CLS, MODE, Pt, T>;
    JFEM_TRY(ensure_dynamic_smem(h, (const void *)kern, smem));
    int per_sm = 1;
    JFEM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, T, smem));
    if (per_sm < 1) per_sm = 1;
    h->last_smem = (int64_t)smem; h->last_blocks_per_sm = per_sm;
    int grid = h->n_sms * per_sm;
    if (grid > D.n_patches) grid = D.n_patches;
    if (a.tail) {
        a.gbar_target = (h->gbar_count += (unsigned)grid);
        Pt ptc = pt;
        void *args[] = {(void *)&a, (void *)&ptc};
        JFEM_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3(grid), dim3(T), args, smem, h->stream));
    } else
        kern<<<grid, T, smem, h->stream>>>(a, pt);
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
    if (h->mat_kind == JFEM_MAT_STVK) {
        if (mode == OP_RESIDUAL) { PtStVKResidual pt; fill_material(h, pt); pt.geo = 0; return launch_set<NNPE, CLASS_GENERAL, OP_RESIDUAL, PtStVKResidual, T>(h, D, a, pt); }
        PtStVKTangent pt; fill_material(h, pt); pt.geo = h->geometric_stiffness ? 1 : 0;
        return launch_set<NNPE, CLASS_GENERAL, OP_TANGENT, PtStVKTangent, T>(h, D, a, pt);
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

// true when op_apply will run exactly one warp-specialised patch kernel (which can carry the halo exchange)
bool ws_halo_capable(jfem_handle *h) {
    if (ensure_built(h) != JFEM_OK) return false;
    if ((h->mesh.nnpe != 10 && h->mesh.nnpe != 8) || h->mat_kind != JFEM_MAT_LINEAR_ELASTIC || !h->warp_specialised || h->patch_elems != WS_T) return false;
    const PatchSetDev &D = h->dsets[CLASS_AFFINE];
    if (h->dsets[CLASS_GENERAL].n_elems != 0 || D.n_elems == 0) return false;
    WsSmem L;
    return ws_layout(D, D.nxr == 0 ? 1 : 0, L);
}

// y = A(x): mode OP_LINEAR (K x), OP_RESIDUAL (f_int(x)), OP_TANGENT (K(ulin) x)
int op_apply(jfem_handle *h, int mode, const double *x, double *y, int flags, const int *done) {
    JFEM_TRY(ensure_built(h));
    if (h->mat_kind < 0) { jfem_set_error("jfem_set_material has not been called"); return JFEM_ESTATE; }
    if (mode == OP_TANGENT && h->mat_kind != JFEM_MAT_LINEAR_ELASTIC && !h->has_lin) {
        jfem_set_error("tangent operator needs jfem_set_linearization"); return JFEM_ESTATE;
    }
    h->matvec_launches = 0;
    const double *x_in = x;
    if ((uintptr_t)x & 15) {   // the 128-bit gathers need a 16-byte aligned operand: take an aligned copy of a misaligned one
        if (h->xal.n != (size_t)h->n_dofs()) JFEM_TRY(h->xal.alloc((size_t)h->n_dofs()));
        JFEM_CUDA(cudaMemcpyAsync(h->xal.p, x, (size_t)h->n_dofs() * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        x = h->xal.p;
    }
    const int atomic_iface = h->deterministic ? 0 : 1;
    const long long n3 = 3LL * (long long)h->inodes.n;
    if (atomic_iface && n3) {   // interface nodes are accumulated with atomicAdd
        nodes_zero_kernel<<<(unsigned)((n3 + 255) / 256), 256, 0, h->stream>>>(h->inodes.p, y, n3, done);
        h->matvec_launches++;
    }
    int last_set = -1;
    for (int c = 0; c < N_CLASSES; c++) if (h->dsets[c].n_elems) last_set = c;
    const bool fused = !atomic_iface && n3 && h->fused_iface && h->coop_ok;
    for (int c = 0; c < N_CLASSES; c++) {
        const PatchSetDev &D = h->dsets[c];
        if (D.n_elems == 0) continue;
        const PatchLayout &L = D.L;
        PatchKArgs a;
        a.tail = (fused && c == last_set) ? 1 : 0;   // the last patch kernel of the operator also reduces the interface nodes
        a.gbar = h->gbar.p; a.gbar_target = 0; a.inodes = h->inodes.p; a.ibase = h->ibase.p; a.n_inodes3 = n3;
        a.halo = 0; a.n_owned = 0xFFFFFFFFu;
        if (h->halo_armed) {
            if (x_in != h->halo_x) { jfem_set_error("internal: fused halo armed for another vector"); return JFEM_ESTATE; }
            a.halo = 1; a.n_owned = (uint32_t)h->n_owned_nodes;
            halo_fill_fused(h, a.hf);
            h->halo_armed = false;
            h->halo_in_kernel = true;
        }
        a.blob = D.blob.p; a.n_patches = D.n_patches; a.stride = L.stride;
        a.offB = L.offB; a.offC = L.offC; a.bytesA = L.bytesA(); a.bytesB = L.bytesB(); a.bytesC = L.bytesC();
        a.a_cx = L.off_cx - L.offA; a.a_cX = L.off_cX - L.offA; a.b_et = L.off_et - L.offB;
        a.c_qn = L.off_qn - L.offC; a.c_ql = L.off_ql - L.offC; a.c_jo = L.off_jo - L.offC;
        a.max_nodes = D.max_nodes; a.max_entries = D.max_entries; a.x_all = D.nxr == 0 ? 1 : 0;
        a.xs_cap = 2 * D.max_ncx; a.Xs_cap = 2 * (a.x_all ? D.max_ncx : D.max_ncX); a.n_dofs = h->n_dofs();
        a.coords = h->coords.p; a.x = x; a.ulin = h->ulin.p; a.y = y; a.ipart = h->ipart.p; a.slot_node = h->slot_node.p;
        a.elem_offset = D.elem_offset;
        a.project = (flags & JFEM_PROJECT) ? 1 : 0; a.atomic_iface = atomic_iface;
        a.fail = h->dflags.p; a.done = done; a.timing = h->timing.p; a.nbuf = 2; a.dbg = h->debug_skip;
        int rc;
        switch (h->mesh.nnpe) {
            case 10: rc = dispatch_threads<10>(h, D, a, mode); break;
            case 8: rc = dispatch_threads<8>(h, D, a, mode); break;
            default: rc = dispatch_threads<4>(h, D, a, mode); break;
        }
        JFEM_TRY(rc);
    }
    if (!atomic_iface && n3 && !fused) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((n3 / 3 + 127) / 128)); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = h->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        JFEM_CUDA(cudaLaunchKernelEx(&cfg, iface_reduce_kernel, (const uint32_t *)h->inodes.p, (const int32_t *)h->ibase.p, (const double *)h->ipart.p, y, n3,
                                     done));
        JFEM_CUDA(cudaGetLastError());
        h->matvec_launches++;
    }
    h->total_launches += h->matvec_launches;
    return JFEM_OK;
}

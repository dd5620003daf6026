// patches.cpp -- host-side construction of the thread-block "patches" the matrix-free operator runs on.
//
// Elements are clustered spatially (recursive coordinate bisection of element centroids, leaves of
// exactly EP elements) so that one thread block owns EP elements and the ~EP*nnpe/valence nodes they
// touch.  For every patch we store one metadata blob (layout in common.h): the node lists, the element table
// (per element node: position of the node in the gathered x tile and the entry of the jagged staging tile that
// receives the element's contribution) and the per-node reduce tables.  The staging tile lets the kernel reduce
// element contributions per node in a FIXED order with no atomics (the deterministic replacement for
// the 30 CUDA.@atomic adds of demos/gpu_assembly_tet10.jl:225-229 and the owner-computes gather of
// ext/JuliaFEMCUDAExt.jl:293-361).  Interface nodes (touched by more than one patch) get one partial-sum slot per
// touching patch, contiguous per node; the last patch to arrive adds them in ascending patch order.
// Elements are assigned to lanes so that the 16 lanes of a half-warp hit distinct shared-memory banks as often as
// possible (greedy, windowed): the random 64-bit gathers / scatters of the element threads are the kernel's main cost.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <numeric>
#include <cstring>
#include <cstdlib>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "common.h"

namespace {

struct Rcb {
    const double *cx;  // centroids 3*n
    std::vector<int64_t> &idx;
    int EP;
    void split(int64_t lo, int64_t hi) {
        int64_t n = hi - lo;
        if (n <= EP) return;
        int64_t leaves = (n + EP - 1) / EP;
        int64_t left = (leaves / 2) * EP;
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (int64_t i = lo; i < hi; i++)
            for (int d = 0; d < 3; d++) {
                double v = cx[3 * idx[i] + d];
                mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v);
            }
        int ax = 0;
        for (int d = 1; d < 3; d++) if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
        const double *c = cx;
        std::nth_element(idx.begin() + lo, idx.begin() + lo + left, idx.begin() + hi, [c, ax](int64_t a, int64_t b) {
            double va = c[3 * a + ax], vb = c[3 * b + ax];
            return va < vb || (va == vb && a < b);
        });
        if (n > (1 << 16)) {
#pragma omp task
            split(lo, lo + left);
#pragma omp task
            split(lo + left, hi);
#pragma omp taskwait
        } else {
            split(lo, lo + left);
            split(lo + left, hi);
        }
    }
};

}  // namespace

// An element is "affine" when its geometry map is affine: Tet10 with every mid-edge node at the midpoint of
// its edge (same idea as the is_CM test of src/assembly/assembly.jl:151-159); Tet4 always.
void classify_elements(MeshHost &m, bool use_affine) {
    m.cls.assign(m.n_elems, CLASS_GENERAL);
    if (!use_affine) return;
    if (m.nnpe == 4) { std::fill(m.cls.begin(), m.cls.end(), (uint8_t)CLASS_AFFINE); return; }
    if (m.nnpe == 8) {   // Hex8 is affine when it is a parallelepiped: x_k = x_0 + (edge vectors of node 0) for every node
        static const int EU[8] = {0, 1, 1, 0, 0, 1, 1, 0}, EV[8] = {0, 0, 1, 1, 0, 0, 1, 1}, EW[8] = {0, 0, 0, 0, 1, 1, 1, 1};
#pragma omp parallel for schedule(static)
        for (int64_t e = 0; e < m.n_elems; e++) {
            const int32_t *c = &m.conn[e * 8];
            const double *x0 = &m.coords[3 * (int64_t)c[0]], *x1 = &m.coords[3 * (int64_t)c[1]], *x3 = &m.coords[3 * (int64_t)c[3]],
                         *x4 = &m.coords[3 * (int64_t)c[4]];
            double h2 = 0;
            for (int d = 0; d < 3; d++) h2 += (x1[d] - x0[d]) * (x1[d] - x0[d]) + (x3[d] - x0[d]) * (x3[d] - x0[d]) + (x4[d] - x0[d]) * (x4[d] - x0[d]);
            bool ok = true;
            for (int k = 0; k < 8 && ok; k++)
                for (int d = 0; d < 3; d++) {
                    const double want = x0[d] + EU[k] * (x1[d] - x0[d]) + EV[k] * (x3[d] - x0[d]) + EW[k] * (x4[d] - x0[d]);
                    const double dv = m.coords[3 * (int64_t)c[k] + d] - want;
                    if (dv * dv > 1e-28 * h2) { ok = false; break; }
                }
            m.cls[e] = ok ? CLASS_AFFINE : CLASS_GENERAL;
        }
        return;
    }
    if (m.nnpe != 10) return;
    static const int EA[6] = {0, 1, 0, 0, 1, 2}, EB[6] = {1, 2, 2, 3, 3, 3};
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < m.n_elems; e++) {
        const int32_t *c = &m.conn[e * 10];
        double h2 = 0;
        for (int d = 0; d < 3; d++) {
            double t = m.coords[3 * (int64_t)c[1] + d] - m.coords[3 * (int64_t)c[0] + d];
            h2 += t * t;
        }
        bool ok = true;
        for (int k = 0; k < 6 && ok; k++)
            for (int d = 0; d < 3; d++) {
                double mid = 0.5 * (m.coords[3 * (int64_t)c[EA[k]] + d] + m.coords[3 * (int64_t)c[EB[k]] + d]);
                double dv = m.coords[3 * (int64_t)c[4 + k] + d] - mid;
                if (dv * dv > 1e-28 * h2) { ok = false; break; }  // |dev| <= 1e-14 * edge length
            }
        m.cls[e] = ok ? CLASS_AFFINE : CLASS_GENERAL;
    }
}

namespace {

// Per-patch work arrays of the lane assignment.  Access "slots" of an element thread: nnpe gathers of x (index = position
// of the node in the x tile), nvx gathers of coordinates (index = coordinate slot) and nnpe scatters into the staging tile
// (index = entry).  64-bit shared-memory accesses are served per half-warp; two lanes conflict when they address different
// words of the same 8-byte bank (word index mod 16; the AoS factor 3 is a bijection mod 16).
struct LaneOpt {
    int ns = 0, ns_assign = 0;        // slots per element; leading slots the lane assignment looks at
    std::vector<uint16_t> idx;        // [elem][slot]
    int model(const std::vector<int> &perm, int ne) const {   // modelled wavefronts (per component) of one patch
        int total = 0;
        for (int h0 = 0; h0 < ne; h0 += 16) {
            int h1 = std::min(h0 + 16, ne);
            for (int s = 0; s < ns; s++) {
                uint16_t seen[16][16];
                int nseen[16] = {0};
                int mx = 1;
                for (int t = h0; t < h1; t++) {
                    uint16_t v = idx[(size_t)perm[t] * ns + s];
                    int b = v & 15, k = 0;
                    for (; k < nseen[b]; k++) if (seen[b][k] == v) break;
                    if (k == nseen[b]) { seen[b][nseen[b]++] = v; mx = std::max(mx, nseen[b]); }
                }
                total += mx;
            }
        }
        return total;
    }
    // Greedy: fill one half-warp at a time; every lane takes, among the next `window` unassigned elements (caller order,
    // so neighbours that share nodes -- broadcasts -- stay together), the one that raises the half-warp's wavefront
    // count (sum over slots of the fullest bank) the least; ties go to fewer bank collisions, then to caller order.
    void assign(std::vector<int> &perm, int ne, int window) const {
        std::vector<uint8_t> used(ne, 0);
        perm.resize(ne);
        int first = 0, lane = 0;
        std::vector<uint16_t> addr((size_t)ns * 16 * 16);   // distinct words per (slot, bank)
        std::vector<uint8_t> cnt((size_t)ns * 16), mx(ns);
        while (lane < ne) {
            std::fill(cnt.begin(), cnt.end(), 0);
            std::fill(mx.begin(), mx.end(), 0);
            for (int s16 = 0; s16 < 16 && lane < ne; s16++) {
                while (used[first]) first++;
                int best = -1, bestcost = 1 << 30, seen = 0;
                for (int i = first; i < ne && seen < window; i++) {
                    if (used[i]) continue;
                    seen++;
                    const uint16_t *v = &idx[(size_t)i * ns];
                    int cost = 0;
                    for (int s = 0; s < ns_assign; s++) {
                        const int sb = s * 16 + (v[s] & 15), n = cnt[sb];
                        const uint16_t *ad = &addr[(size_t)sb * 16];
                        int k = 0;
                        for (; k < n; k++) if (ad[k] == v[s]) break;
                        if (k == n && n > 0) cost += (n + 1 > mx[s]) ? 65 : 1;
                    }
                    if (cost < bestcost) { bestcost = cost; best = i; if (cost == 0) break; }
                }
                used[best] = 1;
                perm[lane++] = best;
                const uint16_t *v = &idx[(size_t)best * ns];
                for (int s = 0; s < ns_assign; s++) {
                    const int sb = s * 16 + (v[s] & 15), n = cnt[sb];
                    uint16_t *ad = &addr[(size_t)sb * 16];
                    int k = 0;
                    for (; k < n; k++) if (ad[k] == v[s]) break;
                    if (k == n) { ad[n] = v[s]; cnt[sb] = (uint8_t)(n + 1); if (n + 1 > mx[s]) mx[s] = (uint8_t)(n + 1); }
                }
            }
        }
    }
};

}  // namespace

int build_patch_sets(const MeshHost &m, int EP, bool use_affine, int lane_window, int64_t n_owned, PatchSetHost sets[N_CLASSES], InterfaceHost &iface) {
    (void)use_affine;
    const int nnpe = m.nnpe;
    if (EP * nnpe > 65535) { jfem_set_error("patch_elems=%d too large for 16-bit staging entries", EP); return JFEM_EINVAL; }
    // centroids
    std::vector<double> cen(3 * (size_t)m.n_elems);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < m.n_elems; e++) {
        double s[3] = {0, 0, 0};
        int nv = (nnpe == 10) ? 4 : nnpe;
        for (int k = 0; k < nv; k++)
            for (int d = 0; d < 3; d++) s[d] += m.coords[3 * (int64_t)m.conn[e * nnpe + k] + d];
        for (int d = 0; d < 3; d++) cen[3 * e + d] = s[d] / nv;
    }
    // Snap the centroids to a grid of about one "cell" (the volume of 6 tets / 1 hex, a Kuhn cell for the lattice meshes):
    // elements of the same cell then compare equal along every axis and the bisection -- ties go by element id -- takes
    // whole cells instead of slicing every cell of the boundary layer, which would put that layer's nodes into both
    // patches (T1: 605 -> ~560 nodes per patch, fewer interface nodes).  Any grid gives a valid clustering.
    if (m.n_elems > 0) {
        double vol = 0;
#pragma omp parallel for schedule(static) reduction(+ : vol)
        for (int64_t e = 0; e < m.n_elems; e++) {
            const int32_t *c = &m.conn[e * nnpe];
            const int i1 = 1, i2 = nnpe == 8 ? 3 : 2, i3 = nnpe == 8 ? 4 : 3;
            double a[3], b[3], d[3];
            for (int k = 0; k < 3; k++) {
                a[k] = m.coords[3 * (int64_t)c[i1] + k] - m.coords[3 * (int64_t)c[0] + k];
                b[k] = m.coords[3 * (int64_t)c[i2] + k] - m.coords[3 * (int64_t)c[0] + k];
                d[k] = m.coords[3 * (int64_t)c[i3] + k] - m.coords[3 * (int64_t)c[0] + k];
            }
            vol += std::fabs(a[0] * (b[1] * d[2] - b[2] * d[1]) - a[1] * (b[0] * d[2] - b[2] * d[0]) + a[2] * (b[0] * d[1] - b[1] * d[0]));
        }
        const double q = std::cbrt(vol / (double)m.n_elems);   // |det| = 6 V for a tet, V for a hex: the cell edge either way
        if (q > 0 && std::isfinite(q)) {
            double org[3] = {1e300, 1e300, 1e300};
            for (int64_t n = 0; n < m.n_nodes; n++)
                for (int k = 0; k < 3; k++) org[k] = std::min(org[k], m.coords[3 * n + k]);
#pragma omp parallel for schedule(static)
            for (int64_t e = 0; e < m.n_elems; e++)
                for (int k = 0; k < 3; k++) cen[3 * e + k] = std::floor((cen[3 * e + k] - org[k]) / q + 1e-6);
        }
    }
    // ---- pass 1: cluster the elements of each class into patches; unique node list (ascending id) per patch
    std::vector<int32_t> touch(m.n_nodes, 0);
    std::vector<std::vector<int32_t>> pn[N_CLASSES];
    for (int c = 0; c < N_CLASSES; c++) {
        PatchSetHost &S = sets[c];
        S = PatchSetHost();
        S.cls = c; S.nnpe = nnpe; S.EP = EP;
        S.nxr = (c == CLASS_AFFINE && (nnpe == 10 || nnpe == 8)) ? 2 : 0;
        for (int64_t e = 0; e < m.n_elems; e++) if (m.cls[e] == c) S.elem_perm.push_back(e);
        S.n_elems = (int64_t)S.elem_perm.size();
        if (S.n_elems == 0) continue;
        Rcb r{cen.data(), S.elem_perm, EP};
#pragma omp parallel
#pragma omp single
        r.split(0, S.n_elems);
        S.n_patches = (int)((S.n_elems + EP - 1) / EP);
        pn[c].resize(S.n_patches);
#pragma omp parallel for schedule(dynamic, 64)
        for (int p = 0; p < S.n_patches; p++) {
            int64_t lo = (int64_t)p * EP, hi = std::min(lo + EP, S.n_elems);
            std::sort(S.elem_perm.begin() + lo, S.elem_perm.begin() + hi);   // caller order inside each patch (deterministic)
            std::vector<int32_t> &v = pn[c][p];
            v.reserve((hi - lo) * nnpe);
            for (int64_t i = lo; i < hi; i++)
                for (int k = 0; k < nnpe; k++) v.push_back(m.conn[S.elem_perm[i] * nnpe + k]);
            std::sort(v.begin(), v.end());
            v.erase(std::unique(v.begin(), v.end()), v.end());
        }
        // Partitioned meshes (nodes >= n_owned are ghosts whose values arrive by the halo exchange): patches that read no
        // ghost value come first, so that the kernel can start on them while the halo is still in flight.  The partial
        // patch (if any) stays last: the internal element order is patch-major with full patches.
        S.ghosty.assign(S.n_patches, 0);
        if (n_owned >= 0 && n_owned < m.n_nodes) {
            for (int p = 0; p < S.n_patches; p++) S.ghosty[p] = pn[c][p].back() >= n_owned ? 1 : 0;   // ids ascending
            const int nfull = (int)(S.n_elems / EP);
            std::vector<int> ord(nfull);
            std::iota(ord.begin(), ord.end(), 0);
            std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return S.ghosty[a] < S.ghosty[b]; });
            std::vector<int64_t> ep2(S.elem_perm.size());
            std::vector<std::vector<int32_t>> pn2(S.n_patches);
            std::vector<uint8_t> gh2(S.n_patches, 0);
            for (int q = 0; q < S.n_patches; q++) {
                const int p = q < nfull ? ord[q] : q;
                const int64_t lo = (int64_t)p * EP, hi = std::min(lo + EP, S.n_elems);
                std::copy(S.elem_perm.begin() + lo, S.elem_perm.begin() + hi, ep2.begin() + (int64_t)q * EP);
                pn2[q].swap(pn[c][p]);
                gh2[q] = S.ghosty[p];
            }
            S.elem_perm.swap(ep2); pn[c].swap(pn2); S.ghosty.swap(gh2);
        }
        S.pnode_ptr.assign(S.n_patches + 1, 0);
        for (int p = 0; p < S.n_patches; p++) {
            if ((int)pn[c][p].size() > 65534) { jfem_set_error("patch has too many nodes"); return JFEM_EINVAL; }
            S.pnode_ptr[p + 1] = S.pnode_ptr[p] + (int32_t)pn[c][p].size();
            S.max_nodes = std::max(S.max_nodes, (int)pn[c][p].size());
            for (int32_t n : pn[c][p]) touch[n]++;
        }
    }
    // ---- interface nodes: contiguous partial slots per node; rank of every (patch, node) among the node's patches
    iface = InterfaceHost();
    std::vector<int32_t> ibase_of(m.n_nodes, -1);
    {
        int64_t tot = 0;
        for (int64_t n = 0; n < m.n_nodes; n++) {
            if (touch[n] > 255) { jfem_set_error("node %lld is shared by more than 255 patches", (long long)n); return JFEM_EINVAL; }
            if (touch[n] > 1) {
                iface.inodes.push_back((uint32_t)n);
                iface.ibase.push_back((int32_t)tot);
                ibase_of[n] = (int32_t)tot;
                tot += touch[n];
                if (tot > (int64_t)PN_ID_MASK) { jfem_set_error("too many interface partial slots"); return JFEM_EINVAL; }
            } else if (touch[n] == 0) iface.orphans.push_back((uint32_t)n);
        }
        iface.n_partials = tot;
    }
    std::vector<uint8_t> prank[N_CLASSES];
    {
        std::vector<uint8_t> seen(m.n_nodes, 0);
        for (int c = 0; c < N_CLASSES; c++) {
            PatchSetHost &S = sets[c];
            if (S.n_elems == 0) continue;
            prank[c].resize(S.pnode_ptr[S.n_patches]);
            for (int p = 0; p < S.n_patches; p++)
                for (size_t j = 0; j < pn[c][p].size(); j++) prank[c][S.pnode_ptr[p] + j] = seen[pn[c][p][j]]++;
        }
    }
    // ---- pass 2: per-patch tables and blobs
    int overflow = 0;
    for (int c = 0; c < N_CLASSES; c++) {
        PatchSetHost &S = sets[c];
        if (S.n_elems == 0) continue;
        const int nvx = S.nxr ? 4 : 0;
        // element nodes whose coordinates an affine element needs: Tet10 vertices; Hex8 node 0 and its three edge neighbours
        static const int XN10[4] = {0, 1, 2, 3}, XN8[4] = {0, 1, 3, 4};
        const int *xn = nnpe == 8 ? XN8 : XN10;
        auto xidx = [&](int k) { for (int v = 0; v < nvx; v++) if (xn[v] == k) return v; return -1; };
        // sizes needed for the layout: chunk counts of the gather lists, max_rows
        std::vector<int> ncxs(S.n_patches, 0), ncXs(S.n_patches, 0), nrw(S.n_patches, 0);
#pragma omp parallel for schedule(dynamic, 64)
        for (int p = 0; p < S.n_patches; p++) {
            int64_t lo = (int64_t)p * EP, hi = std::min(lo + EP, S.n_elems);
            const std::vector<int32_t> &ids = pn[c][p];
            std::vector<int32_t> cnt(ids.size(), 0);
            std::vector<uint8_t> nx(ids.size(), 0);
            for (int64_t i = lo; i < hi; i++)
                for (int k = 0; k < nnpe; k++) {
                    int j = (int)(std::lower_bound(ids.begin(), ids.end(), m.conn[S.elem_perm[i] * nnpe + k]) - ids.begin());
                    cnt[j]++;
                    if (xidx(k) >= 0) nx[j] = 1;
                }
            nrw[p] = *std::max_element(cnt.begin(), cnt.end());
            // every node occupies the two chunks floor(3n/2), floor(3n/2)+1; ids ascending
            int64_t last = -1, lastX = -1;
            int nc = 0, ncX = 0;
            for (size_t j = 0; j < ids.size(); j++) {
                const int64_t c0 = (3 * (int64_t)ids[j]) >> 1;
                nc += (c0 > last) + (c0 + 1 > last); last = c0 + 1;
                if (nx[j]) { ncX += (c0 > lastX) + (c0 + 1 > lastX); lastX = c0 + 1; }
            }
            ncxs[p] = nc; ncXs[p] = ncX;
        }
        for (int p = 0; p < S.n_patches; p++) {
            S.max_ncx = std::max(S.max_ncx, ncxs[p]); S.max_ncX = std::max(S.max_ncX, ncXs[p]); S.max_rows = std::max(S.max_rows, nrw[p]);
        }
        if (S.max_ncx > 32767) { jfem_set_error("patch needs too many gather chunks"); return JFEM_EINVAL; }
        if (S.max_rows > 255) { jfem_set_error("a node has more than 255 elements inside one patch"); return JFEM_EINVAL; }
        S.max_entries = EP * nnpe;
        auto r16 = [](int v) { return (v + 15) & ~15; };
        PatchLayout &L = S.L;
        L.offA = 0; L.off_cx = 16; L.off_cX = L.off_cx + r16(4 * S.max_ncx);
        L.offB = L.off_cX + r16(4 * S.max_ncX); L.off_et = L.offB + 16;
        L.offC = L.off_et + r16(4 * (nnpe + S.nxr) * EP);
        L.off_qn = L.offC + 16; L.off_ql = L.off_qn + r16(4 * S.max_nodes);
        L.off_jo = L.off_ql + r16(S.max_nodes);
        L.stride = L.off_jo + r16(2 * (S.max_rows + 1));
        S.blob.assign((size_t)L.stride * S.n_patches, 0);
        S.qnodes.assign((size_t)S.pnode_ptr[S.n_patches], 0);
        S.qids.assign((size_t)S.pnode_ptr[S.n_patches], 0);
        std::vector<int64_t> new_perm(S.elem_perm.size());
        double wf0 = 0, wf1 = 0, wfi = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : wf0, wf1, wfi) reduction(| : overflow)
        for (int p = 0; p < S.n_patches; p++) {
            const int64_t lo = (int64_t)p * EP, hi = std::min(lo + EP, S.n_elems);
            const int ne = (int)(hi - lo);
            const std::vector<int32_t> &ids = pn[c][p];
            const int np = (int)ids.size(), nb = S.pnode_ptr[p];
            auto local_of = [&](int32_t n) { return (int)(std::lower_bound(ids.begin(), ids.end(), n) - ids.begin()); };
            std::vector<int32_t> cnt(np, 0);
            std::vector<int32_t> xslot(np, -1);      // >= 0: the node's coordinates are needed; later its offset in the coordinate tile
            std::vector<int32_t> xoff(np, 0);        // offset (in doubles) of the node in the x tile
            std::vector<uint32_t> chx, chX;          // gather chunk lists
            std::vector<uint16_t> loc((size_t)ne * nnpe);
            for (int i = 0; i < ne; i++)
                for (int k = 0; k < nnpe; k++) {
                    int j = local_of(m.conn[S.elem_perm[lo + i] * nnpe + k]);
                    loc[(size_t)i * nnpe + k] = (uint16_t)j;
                    cnt[j]++;
                    if (xidx(k) >= 0) xslot[j] = 0;
                }
            // gather chunks: node n = doubles 3n..3n+2 = chunks floor(3n/2), +1; its offset in the tile = 2 * (index of its
            // first chunk) + (3n & 1)
            for (int j = 0; j < np; j++) {
                const int64_t d0 = 3 * (int64_t)ids[j], c0 = d0 >> 1;
                if (chx.empty() || (int64_t)chx.back() < c0) chx.push_back((uint32_t)c0);
                const int first = (int)chx.size() - 1 - ((int64_t)chx.back() > c0 ? 1 : 0);
                if ((int64_t)chx.back() < c0 + 1) chx.push_back((uint32_t)(c0 + 1));
                xoff[j] = 2 * first + (int)(d0 & 1);
                if (xslot[j] == 0) {
                    if (chX.empty() || (int64_t)chX.back() < c0) chX.push_back((uint32_t)c0);
                    const int fX = (int)chX.size() - 1 - ((int64_t)chX.back() > c0 ? 1 : 0);
                    if ((int64_t)chX.back() < c0 + 1) chX.push_back((uint32_t)(c0 + 1));
                    xslot[j] = 2 * fX + (int)(d0 & 1);
                }
            }
            const int ncx = (int)chx.size(), ncX = (int)chX.size();
            // reduce order q: descending contribution count, ascending id inside a count
            std::vector<int> order(np), qpos(np);
            std::iota(order.begin(), order.end(), 0);
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cnt[a] > cnt[b]; });
            for (int q = 0; q < np; q++) qpos[order[q]] = q;
            const int nrows = cnt[order[0]];
            std::vector<int> joff(nrows + 1, 0);
            for (int r = 0; r < nrows; r++) {
                int w = 0;
                while (w < np && cnt[order[w]] > r) w++;   // np*nrows is small
                joff[r + 1] = joff[r] + w;
            }
            // ---- lane assignment (on the gather slots), then the staging rows.  Slot order: x gathers | coordinate gathers | scatters
            LaneOpt lo_;
            lo_.ns = 2 * nnpe + nvx;
            lo_.ns_assign = nnpe + nvx;
            lo_.idx.resize((size_t)ne * lo_.ns);
            std::vector<uint16_t> ent((size_t)ne * nnpe);
            {
                std::vector<int> fill(np, 0);   // baseline: rank = caller order of the element inside the patch
                for (int i = 0; i < ne; i++)
                    for (int k = 0; k < nnpe; k++) {
                        const int j = loc[(size_t)i * nnpe + k];
                        const int e = joff[fill[j]++] + qpos[j];
                        ent[(size_t)i * nnpe + k] = (uint16_t)e;
                        lo_.idx[(size_t)i * lo_.ns + k] = (uint16_t)xoff[j];
                        if (xidx(k) >= 0) lo_.idx[(size_t)i * lo_.ns + nnpe + xidx(k)] = (uint16_t)xslot[j];
                        lo_.idx[(size_t)i * lo_.ns + nnpe + nvx + k] = (uint16_t)e;
                    }
            }
            std::vector<int> perm(ne);
            std::iota(perm.begin(), perm.end(), 0);
            wf0 += lo_.model(perm, ne);
            if (lane_window > 0) {
                lo_.assign(perm, ne, lane_window);
                // Staging rows: a node's contributions may take its rows in any order (the order only fixes the summation
                // order).  Per half-warp and element node k, give every lane a free row of its node whose bank is still
                // unused in that store instruction; most constrained nodes first.
                std::vector<int> rstart(np + 1, 0);
                for (int j = 0; j < np; j++) rstart[j + 1] = rstart[j] + cnt[j];
                std::vector<uint8_t> rused(rstart[np], 0);
                for (int h0 = 0; h0 < ne; h0 += 16) {
                    const int h1 = std::min(h0 + 16, ne);
                    for (int k = 0; k < nnpe; k++) {
                        // bipartite matching lanes -> banks (augmenting paths); a lane may use bank b if its node has a free row there
                        const int nl = h1 - h0;
                        int jn[16], rowof[16][16], owner[16], got[16];
                        for (int q = 0; q < nl; q++) {
                            jn[q] = loc[(size_t)perm[h0 + q] * nnpe + k];
                            got[q] = -1;
                            for (int bb = 0; bb < 16; bb++) rowof[q][bb] = -1;
                        }
                        for (int bb = 0; bb < 16; bb++) owner[bb] = -1;
                        // lanes sharing a node compete for its rows: hand each free row to one lane only (round-robin)
                        for (int q = 0; q < nl; q++) {
                            const int j = jn[q];
                            int first = q, nshare = 0, myidx = 0;
                            for (int q2 = 0; q2 < nl; q2++) if (jn[q2] == j) { if (q2 < first) first = q2; if (q2 < q) myidx++; nshare++; }
                            int f = 0;
                            for (int r = 0; r < cnt[j]; r++) {
                                if (rused[rstart[j] + r]) continue;
                                if (f++ % nshare != myidx) continue;
                                const int bb = (joff[r] + qpos[j]) & 15;
                                if (rowof[q][bb] < 0) rowof[q][bb] = r;
                            }
                        }
                        for (int q = 0; q < nl; q++) {
                            bool vis[16] = {false};
                            struct Aug {
                                static bool go(int q, int (*rowof)[16], int *owner, int *got, bool *vis) {
                                    for (int bb = 0; bb < 16; bb++) {
                                        if (rowof[q][bb] < 0 || vis[bb]) continue;
                                        vis[bb] = true;
                                        if (owner[bb] < 0 || go(owner[bb], rowof, owner, got, vis)) { owner[bb] = q; got[q] = bb; return true; }
                                    }
                                    return false;
                                }
                            };
                            Aug::go(q, rowof, owner, got, vis);
                        }
                        int bank[16] = {0};
                        for (int q = 0; q < nl; q++) if (got[q] >= 0) bank[got[q]]++;
                        for (int pass = 0; pass < 2; pass++)   // matched lanes first, so that the others cannot take their rows
                            for (int q = 0; q < nl; q++) {
                                if ((got[q] >= 0) != (pass == 0)) continue;
                                const int i = perm[h0 + q], j = jn[q];
                                int br = -1;
                                if (got[q] >= 0) br = rowof[q][got[q]];
                                else {   // unmatched: least loaded bank among the node's free rows
                                    int bc = 1 << 30;
                                    for (int r = 0; r < cnt[j]; r++) {
                                        if (rused[rstart[j] + r]) continue;
                                        const int c2 = bank[(joff[r] + qpos[j]) & 15];
                                        if (c2 < bc) { bc = c2; br = r; }
                                    }
                                    bank[(joff[br] + qpos[j]) & 15]++;
                                }
                                rused[rstart[j] + br] = 1;
                                const int e = joff[br] + qpos[j];
                                ent[(size_t)i * nnpe + k] = (uint16_t)e;
                                lo_.idx[(size_t)i * lo_.ns + nnpe + nvx + k] = (uint16_t)e;
                            }
                    }
                }
            }
            wf1 += lo_.model(perm, ne);
            wfi += lo_.ns * ((ne + 15) / 16);
            for (int t = 0; t < ne; t++) new_perm[lo + t] = S.elem_perm[lo + perm[t]];
            // ---- write the blob
            uint8_t *b = &S.blob[(size_t)p * L.stride];
            const int32_t hdr[4] = {np, ncx | (ncX << 16), ne, nrows | (S.ghosty[p] ? 0x10000 : 0)};   // bit 16: the patch reads ghost values
            memcpy(b + L.offA, hdr, 16); memcpy(b + L.offB, hdr, 16); memcpy(b + L.offC, hdr, 16);
            uint32_t *bcx = reinterpret_cast<uint32_t *>(b + L.off_cx), *bcX = reinterpret_cast<uint32_t *>(b + L.off_cX);
            uint32_t *bet = reinterpret_cast<uint32_t *>(b + L.off_et);
            uint32_t *bqn = reinterpret_cast<uint32_t *>(b + L.off_qn);
            uint8_t *bql = b + L.off_ql;
            uint16_t *bjo = reinterpret_cast<uint16_t *>(b + L.off_jo);
            for (int j = 0; j < ncx; j++) bcx[j] = chx[j];
            for (int j = 0; j < ncX; j++) bcX[j] = chX[j];
            for (int t = 0; t < ne; t++) {
                const int i = perm[t];
                for (int k = 0; k < nnpe; k++) bet[(size_t)k * EP + t] = (uint32_t)xoff[loc[(size_t)i * nnpe + k]] | ((uint32_t)ent[(size_t)i * nnpe + k] << 16);
                for (int v = 0; v < nvx; v += 2)
                    bet[(size_t)(nnpe + v / 2) * EP + t] =
                        (uint32_t)xslot[loc[(size_t)i * nnpe + xn[v]]] | ((uint32_t)xslot[loc[(size_t)i * nnpe + xn[v + 1]]] << 16);
            }
            for (int q = 0; q < np; q++) {
                const int j = order[q];
                const int32_t n = ids[j];
                // interior node: its id (the sum is stored to y); interface node: the partial slot of this (patch, node)
                uint32_t w = touch[n] > 1 ? (PN_IFACE | (uint32_t)(ibase_of[n] + (int32_t)prank[c][nb + j])) : (uint32_t)n;
                if (cnt[j] > 255) overflow |= 1;
                bqn[q] = w;
                bql[q] = (uint8_t)cnt[j];
                S.qnodes[nb + q] = w;
                S.qids[nb + q] = n;
            }
            for (int r = 0; r <= nrows; r++) bjo[r] = (uint16_t)joff[r];
        }
        S.elem_perm.swap(new_perm);
        S.wf_before = wf0 / S.n_patches; S.wf_after = wf1 / S.n_patches; S.wf_ideal = wfi / S.n_patches;
    }
    if (overflow) { jfem_set_error("a node has more than 255 elements inside one patch"); return JFEM_EINVAL; }
    return JFEM_OK;
}


// ------------------------------------------------------------------------------------------------ assembled path (host side)

static void node_to_elements(const MeshHost &m, std::vector<int64_t> &nptr, std::vector<int32_t> &n2e) {
    const int nnpe = m.nnpe;
    const int64_t nn = m.n_nodes, ne = m.n_elems;
    nptr.assign(nn + 1, 0);
    for (int64_t i = 0; i < ne * nnpe; i++) nptr[m.conn[i] + 1]++;
    for (int64_t i = 0; i < nn; i++) nptr[i + 1] += nptr[i];
    n2e.resize(nptr[nn]);
    std::vector<int64_t> fill(nptr.begin(), nptr.end() - 1);
    for (int64_t e = 0; e < ne; e++)
        for (int k = 0; k < nnpe; k++) n2e[fill[m.conn[e * nnpe + k]]++] = (int32_t)e;
}

int build_node_adjacency(const MeshHost &m, std::vector<int64_t> &ap, std::vector<int32_t> &adj, std::vector<uint16_t> &eblk) {
    const int nnpe = m.nnpe;
    const int64_t nn = m.n_nodes, ne = m.n_elems;
    std::vector<int64_t> nptr;
    std::vector<int32_t> n2e;
    const bool verbose = getenv("JFEM_SETUP_TIMING") != nullptr;
    double tt[4] = {0, 0, 0, 0};
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    tt[0] = now();
    node_to_elements(m, nptr, n2e);
    tt[1] = now();
    // Every thread takes one contiguous range of nodes, sorts + uniques the candidates of each node in a scratch buffer and
    // appends the result to its own list; the lists are then copied to their place in adj (one sort per node, no per-node
    // allocation).
    ap.assign(nn + 1, 0);
    int nthr = 1;
#ifdef _OPENMP
    nthr = omp_get_max_threads();
#endif
    std::vector<std::vector<int32_t>> tbuf(nthr);
    std::vector<int64_t> tfirst(nthr + 1, nn);
    bool too_many = false;
#pragma omp parallel num_threads(nthr)
    {
        int t = 0, nt = 1;   // the team may be smaller than requested: split by the ACTUAL team size
#ifdef _OPENMP
        t = omp_get_thread_num();
        nt = omp_get_num_threads();
#endif
        const int64_t a0 = nn * t / nt, a1 = nn * (t + 1) / nt;
        tfirst[t] = a0;
        std::vector<int32_t> &out = tbuf[t];
        out.reserve((size_t)((nptr[a1] - nptr[a0]) * 3 + 64));
        // duplicates are dropped through a small open-addressing table before the sort (a Tet10 node sees ~70 candidates
        // for ~28 distinct neighbours); only the slots that were filled are cleared again
        std::vector<int32_t> v, table(256, -1);
        std::vector<uint32_t> slots;
        for (int64_t a = a0; a < a1; a++) {
            v.clear();
            slots.clear();
            const size_t cand = (size_t)(nptr[a + 1] - nptr[a]) * nnpe;
            if (table.size() < 4 * cand) { size_t sz = table.size(); while (sz < 4 * cand) sz *= 2; table.assign(sz, -1); }
            const uint32_t msk = (uint32_t)table.size() - 1;
            for (int64_t q = nptr[a]; q < nptr[a + 1]; q++) {
                const int32_t *c = &m.conn[(int64_t)n2e[q] * nnpe];
                for (int k = 0; k < nnpe; k++) {
                    const int32_t id = c[k];
                    uint32_t hsh = ((uint32_t)id * 2654435761u) & msk;
                    while (table[hsh] >= 0 && table[hsh] != id) hsh = (hsh + 1) & msk;
                    if (table[hsh] < 0) { table[hsh] = id; v.push_back(id); slots.push_back(hsh); }
                }
            }
            for (uint32_t sl : slots) table[sl] = -1;
            std::sort(v.begin(), v.end());
            const size_t nu = v.size();
            if (nu > 65535) too_many = true;
            ap[a + 1] = (int64_t)nu;
            out.insert(out.end(), v.begin(), v.end());
        }
    }
    if (too_many) { jfem_set_error("a node has more than 65535 neighbours"); return JFEM_EINVAL; }
    for (int64_t a = 0; a < nn; a++) ap[a + 1] += ap[a];
    adj.resize(ap[nn]);
#pragma omp parallel for schedule(static, 1) num_threads(nthr)
    for (int t = 0; t < nthr; t++)
        if (!tbuf[t].empty()) std::copy(tbuf[t].begin(), tbuf[t].end(), adj.begin() + ap[tfirst[t]]);
    std::vector<std::vector<int32_t>>().swap(tbuf);
    tt[2] = now();
    // block positions: the nnpe nodes of an element inside the row of each of its nodes
    eblk.resize((size_t)ne * nnpe * nnpe);
#pragma omp parallel for schedule(static)
    for (int64_t e = 0; e < ne; e++)
        for (int k = 0; k < nnpe; k++) {
            const int32_t a = m.conn[e * nnpe + k];
            const int32_t *lo = &adj[ap[a]], *hi = lo + (ap[a + 1] - ap[a]);
            for (int l = 0; l < nnpe; l++) eblk[(e * nnpe + k) * nnpe + l] = (uint16_t)(std::lower_bound(lo, hi, m.conn[e * nnpe + l]) - lo);
        }
    tt[3] = now();
    if (verbose) fprintf(stderr, "[jfem] adjacency: node->elements %.3f s, sort/unique %.3f s, element blocks %.3f s (%d threads)\n", tt[1] - tt[0], tt[2] - tt[1], tt[3] - tt[2], nthr);
    return JFEM_OK;
}

void greedy_colouring(const MeshHost &m, std::vector<int64_t> &colour_ptr, std::vector<int32_t> &celems) {
    const int nnpe = m.nnpe;
    const int64_t ne = m.n_elems;
    std::vector<int64_t> nptr;
    std::vector<int32_t> n2e;
    node_to_elements(m, nptr, n2e);
    // sequential by definition (an element takes the lowest colour none of its already coloured neighbours has); the
    // colours of the neighbours are collected in a 64-bit mask per step, with a byte array only beyond 64 colours
    std::vector<int32_t> colour(ne, -1);
    int ncol = 0;
    std::vector<uint8_t> used;
    for (int64_t e = 0; e < ne; e++) {
        uint64_t mask = 0;
        for (int k = 0; k < nnpe; k++) {
            const int32_t a = m.conn[e * nnpe + k];
            for (int64_t q = nptr[a]; q < nptr[a + 1]; q++) {
                const int32_t c = colour[n2e[q]];
                if (c >= 0 && c < 64) mask |= (1ull << c);
            }
        }
        int c;
        if (~mask) c = __builtin_ctzll(~mask);   // lowest colour below 64 that no neighbour has
        else {                                   // all of 0..63 taken: search the higher colours
            used.assign(ncol + 1, 0);
            for (int k = 0; k < nnpe; k++) {
                const int32_t a = m.conn[e * nnpe + k];
                for (int64_t q = nptr[a]; q < nptr[a + 1]; q++) { const int32_t cc = colour[n2e[q]]; if (cc >= 0) used[cc] = 1; }
            }
            c = 64;
            while (used[c]) c++;                 // used[ncol] == 0: terminates with c <= ncol (a new colour)
        }
        colour[e] = c;
        if (c + 1 > ncol) ncol = c + 1;
    }
    colour_ptr.assign(ncol + 1, 0);
    for (int64_t e = 0; e < ne; e++) colour_ptr[colour[e] + 1]++;
    for (int c = 0; c < ncol; c++) colour_ptr[c + 1] += colour_ptr[c];
    celems.resize(ne);
    std::vector<int64_t> fill(colour_ptr.begin(), colour_ptr.end() - 1);
    for (int64_t e = 0; e < ne; e++) celems[fill[colour[e]]++] = (int32_t)e;
}

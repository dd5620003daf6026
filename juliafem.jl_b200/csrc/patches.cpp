// patches.cpp -- host-side construction of the thread-block "patches" the matrix-free operator runs on.
//
// Elements are clustered spatially (recursive coordinate bisection of element centroids, leaves of
// exactly EP elements) so that one thread block owns EP elements and the ~EP*nnpe/valence nodes they
// touch.  For every patch we store (i) its unique node list (interface nodes -- those touched by more
// than one patch -- first), (ii) the element->local-node table in node-major order (coalesced u16
// reads), (iii) the element->staging-position table (node-major staging: all contributions to one node are
// contiguous), which lets the kernel reduce
// element contributions per node in a FIXED order with no atomics (the deterministic replacement for
// the 30 CUDA.@atomic adds of demos/gpu_assembly_tet10.jl:225-229 and the owner-computes gather of
// ext/JuliaFEMCUDAExt.jl:293-361).  Interface nodes get one partial-sum slot per touching patch; a
// second tiny kernel adds those in ascending patch order.
#include <algorithm>
#include <cmath>
#include <numeric>
#include <cstring>

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

int build_patch_sets(const MeshHost &m, int EP, bool use_affine, PatchSetHost sets[N_CLASSES], InterfaceHost &iface) {
    const int nnpe = m.nnpe;
    if (EP * 3 * nnpe > 65535) { jfem_set_error("patch_elems=%d too large for 16-bit staging slots", EP); return JFEM_EINVAL; }
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
    std::vector<int32_t> touch(m.n_nodes, 0);
    int64_t elem_offset = 0;
    for (int c = 0; c < N_CLASSES; c++) {
        PatchSetHost &S = sets[c];
        S = PatchSetHost();
        S.cls = c; S.nnpe = nnpe; S.EP = EP;
        for (int64_t e = 0; e < m.n_elems; e++) if (m.cls[e] == c) S.elem_perm.push_back(e);
        S.n_elems = (int64_t)S.elem_perm.size();
        if (S.n_elems == 0) continue;
        Rcb r{cen.data(), S.elem_perm, EP};
#pragma omp parallel
#pragma omp single
        r.split(0, S.n_elems);
        S.n_patches = (int)((S.n_elems + EP - 1) / EP);
        // keep caller order inside each patch (stable, deterministic)
#pragma omp parallel for schedule(dynamic, 64)
        for (int p = 0; p < S.n_patches; p++) {
            int64_t lo = (int64_t)p * EP, hi = std::min(lo + EP, S.n_elems);
            std::sort(S.elem_perm.begin() + lo, S.elem_perm.begin() + hi);
        }
        // unique nodes per patch
        S.pnode_ptr.assign(S.n_patches + 1, 0);
        std::vector<std::vector<int32_t>> pn(S.n_patches);
#pragma omp parallel for schedule(dynamic, 64)
        for (int p = 0; p < S.n_patches; p++) {
            int64_t lo = (int64_t)p * EP, hi = std::min(lo + EP, S.n_elems);
            std::vector<int32_t> &v = pn[p];
            v.reserve((hi - lo) * nnpe);
            for (int64_t i = lo; i < hi; i++)
                for (int k = 0; k < nnpe; k++) v.push_back(m.conn[S.elem_perm[i] * nnpe + k]);
            std::sort(v.begin(), v.end());
            v.erase(std::unique(v.begin(), v.end()), v.end());
        }
        for (int p = 0; p < S.n_patches; p++) {
            if ((int)pn[p].size() > 65534) { jfem_set_error("patch has too many nodes"); return JFEM_EINVAL; }
            S.pnode_ptr[p + 1] = S.pnode_ptr[p] + (int32_t)pn[p].size();
            S.max_nodes = std::max(S.max_nodes, (int)pn[p].size());
            for (int32_t n : pn[p]) touch[n]++;
        }
        S.pnodes.resize(S.pnode_ptr[S.n_patches]);
        for (int p = 0; p < S.n_patches; p++) std::copy(pn[p].begin(), pn[p].end(), S.pnodes.begin() + S.pnode_ptr[p]);
        elem_offset += S.n_elems;
    }
    // second pass: order patch nodes (interface first, then by descending incidence count so that the per-node
    // gather loop has warp-uniform trip counts), build local tables
    int64_t ipart_total = 0;
    std::vector<std::vector<int32_t>> node_slots;  // filled below per interface node via sort
    std::vector<std::pair<int32_t, int32_t>> ipairs; // (node, slot) in ascending (set, patch) order
    for (int c = 0; c < N_CLASSES; c++) {
        PatchSetHost &S = sets[c];
        if (S.n_elems == 0) continue;
        S.n_iface.assign(S.n_patches, 0);
        S.ipart_base.assign(S.n_patches, 0);
        S.lconn.assign((size_t)S.n_patches * nnpe * EP, 0xFFFF);
        S.goff.assign((size_t)S.pnode_ptr[S.n_patches] + S.n_patches, 0);
        S.gslots.assign((size_t)S.n_patches * EP * nnpe, 0);
        S.xslot.assign((size_t)S.pnode_ptr[S.n_patches], 0xFFFF);
        std::vector<int> nxs(S.n_patches, 0);
#pragma omp parallel for schedule(dynamic, 64)
        for (int p = 0; p < S.n_patches; p++) {
            int64_t lo = (int64_t)p * EP, hi = std::min(lo + EP, S.n_elems);
            int nb = S.pnode_ptr[p], np = S.pnode_ptr[p + 1] - nb;
            std::vector<int32_t> ids(S.pnodes.begin() + nb, S.pnodes.begin() + nb + np);  // ascending
            std::vector<int32_t> cnt(np, 0);
            std::vector<uint8_t> needx(np, 0);
            auto local_of = [&](int32_t n) { return (int)(std::lower_bound(ids.begin(), ids.end(), n) - ids.begin()); };
            for (int64_t i = lo; i < hi; i++)
                for (int k = 0; k < nnpe; k++) {
                    int j = local_of(m.conn[S.elem_perm[i] * nnpe + k]);
                    cnt[j]++;
                    if (c == CLASS_GENERAL || k < 4) needx[j] = 1;
                }
            std::vector<int> order(np);
            std::iota(order.begin(), order.end(), 0);
            // Order of the patch nodes: interface nodes first (their partial slot is their position), then element
            // VERTEX nodes before mid-side nodes (vertices collect ~3x more contributions, so warps of the per-node
            // reduction get uniform trip counts), ascending id inside each class (locally consecutive addresses).
            std::vector<uint8_t> isvert(np, 0);
            for (int64_t i = lo; i < hi; i++)
                for (int k = 0; k < (nnpe == 10 ? 4 : nnpe); k++) isvert[local_of(m.conn[S.elem_perm[i] * nnpe + k])] = 1;
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
                bool ia = touch[ids[a]] > 1, ib = touch[ids[b]] > 1;
                if (ia != ib) return ia;
                return isvert[a] > isvert[b];
            });
            std::vector<int> newpos(np);
            int nif = 0;
            for (int q = 0; q < np; q++) {
                int j = order[q];
                newpos[j] = q;
                uint32_t w = (uint32_t)ids[j];
                if (needx[j]) w |= PN_NEEDX;
                if (touch[ids[j]] > 1) { w |= PN_IFACE; nif++; }
                S.pnodes[nb + q] = w;
            }
            S.n_iface[p] = nif;
            int nx = 0;
            for (int q = 0; q < np; q++) if (S.pnodes[nb + q] & PN_NEEDX) S.xslot[nb + q] = (uint16_t)nx++;
            nxs[p] = nx;
            // offsets
            uint16_t *go = &S.goff[(size_t)nb + p];
            go[0] = 0;
            for (int q = 0; q < np; q++) go[q + 1] = (uint16_t)(go[q] + cnt[order[q]]);
            std::vector<int> fill(np);
            for (int q = 0; q < np; q++) fill[q] = go[q];
            uint16_t *gs = &S.gslots[(size_t)p * EP * nnpe];
            for (int64_t i = lo; i < hi; i++) {
                int t = (int)(i - lo);
                for (int k = 0; k < nnpe; k++) {
                    int q = newpos[local_of(m.conn[S.elem_perm[i] * nnpe + k])];
                    S.lconn[((size_t)p * nnpe + k) * EP + t] = (uint16_t)q;
                    gs[(size_t)k * EP + t] = (uint16_t)(fill[q]++ - go[q]);   // rank of this element among the node's contributions
                }
            }
        }
        for (int p = 0; p < S.n_patches; p++) S.max_nx = std::max(S.max_nx, nxs[p]);
        for (int p = 0; p < S.n_patches; p++) {
            S.ipart_base[p] = (int32_t)ipart_total;
            int nb = S.pnode_ptr[p];
            for (int q = 0; q < S.n_iface[p]; q++) ipairs.emplace_back((int32_t)(S.pnodes[nb + q] & PN_ID_MASK), (int32_t)(ipart_total + q));
            ipart_total += S.n_iface[p];
            if (ipart_total > 0x7FFFFFF0LL) { jfem_set_error("too many interface partial slots"); return JFEM_EINVAL; }
        }
    }
    // pack the per-patch blobs
    bool rank_overflow = false;
    for (int c = 0; c < N_CLASSES; c++) {
        PatchSetHost &S = sets[c];
        if (S.n_elems == 0) continue;
        auto r16 = [](int v) { return (v + 15) & ~15; };
        S.off_pn = 16;
        S.off_xl = S.off_pn + r16(4 * S.max_nodes);
        S.off_xs = S.off_xl + r16(4 * S.max_nx);
        S.off_go = S.off_xs + r16(2 * S.max_nodes);
        S.off_gs = S.off_go + r16(2 * (S.max_nodes + 1));
        S.off_lc = S.off_gs + r16(EP * nnpe);
        S.stride = S.off_lc + r16(2 * EP * nnpe);
        S.blob.assign((size_t)S.stride * S.n_patches, 0);
#pragma omp parallel for schedule(static)
        for (int p = 0; p < S.n_patches; p++) {
            uint8_t *b = &S.blob[(size_t)p * S.stride];
            const int nb = S.pnode_ptr[p], np = S.pnode_ptr[p + 1] - nb;
            int nx = 0;
            uint32_t *xl = reinterpret_cast<uint32_t *>(b + S.off_xl);
            for (int q = 0; q < np; q++) if (S.pnodes[nb + q] & PN_NEEDX) xl[nx++] = S.pnodes[nb + q] & PN_ID_MASK;
            if ((S.n_iface[p] | nx) > 0xFFFF) nx = 0xFFFF;   // cannot happen: np <= 65534
            int32_t hdr[4] = {np, S.n_iface[p] | (nx << 16), S.ipart_base[p], (int32_t)std::min<int64_t>(EP, S.n_elems - (int64_t)p * EP)};
            memcpy(b, hdr, 16);
            memcpy(b + S.off_pn, &S.pnodes[nb], 4 * (size_t)np);
            memcpy(b + S.off_xs, &S.xslot[nb], 2 * (size_t)np);
            memcpy(b + S.off_go, &S.goff[(size_t)nb + p], 2 * (size_t)(np + 1));
            for (int q = 0; q < EP * nnpe; q++) {
                uint16_t r = S.gslots[(size_t)p * EP * nnpe + q];
                b[S.off_gs + q] = (uint8_t)(r > 255 ? 255 : r);
                if (r > 255) rank_overflow = true;
            }
            memcpy(b + S.off_lc, &S.lconn[(size_t)p * nnpe * EP], 2 * (size_t)EP * nnpe);
        }
        std::vector<uint16_t>().swap(S.lconn);
        std::vector<uint16_t>().swap(S.gslots);
        std::vector<uint16_t>().swap(S.goff);
        std::vector<uint16_t>().swap(S.xslot);
    }
    if (rank_overflow) { jfem_set_error("a node has more than 255 elements inside one patch"); return JFEM_EINVAL; }
    // interface node -> partial slots (ascending slot = ascending (set, patch))
    std::stable_sort(ipairs.begin(), ipairs.end(), [](const std::pair<int32_t, int32_t> &a, const std::pair<int32_t, int32_t> &b) { return a.first < b.first; });
    iface = InterfaceHost();
    iface.n_partials = ipart_total;
    iface.iptr.push_back(0);
    for (size_t i = 0; i < ipairs.size(); i++) {
        if (i == 0 || ipairs[i].first != ipairs[i - 1].first) {
            if (i) iface.iptr.push_back((int32_t)i);
            iface.inodes.push_back((uint32_t)ipairs[i].first);
        }
        iface.islots.push_back(ipairs[i].second);
    }
    if (!ipairs.empty()) iface.iptr.push_back((int32_t)ipairs.size());
    // nodes no element touches: listed with an empty slot range so that the reduce kernel stores y = 0 for them
    for (int64_t n = 0; n < m.n_nodes; n++)
        if (touch[n] == 0) {
            iface.inodes.push_back((uint32_t)n);
            iface.iptr.push_back((int32_t)ipairs.size());
        }
    return JFEM_OK;
}

// hostcheck.cu -- TEST INFRASTRUCTURE (never loaded by the product): runs the patch pipeline of csrc/matvec.cu on the
// CPU, using the very tables patches.cpp builds and the same element code (patch_elem.cuh / elem.cuh compiled for the
// host).  It validates the blob layout (gather order, element table, jagged staging entries, reduce order, interface
// slots) without a GPU; the kernels' synchronisation is of course only exercised by the -m gpu tests.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../juliafem.jl_b200/csrc/patch_elem.cuh"

static char g_err[512];
void jfem_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

using namespace jf;

template <int NNPE, int CLS, int T>
static void run_set(const PatchSetHost &S, long long elem_offset, const MeshHost &m, const PtLinear &pt, const double *x, double *y, int project,
                    std::vector<double> &ipart, long long n_owned, const double *land) {
    const PatchLayout &L = S.L;
    const bool x_all = S.nxr == 0;
    const long long nd = 3 * (long long)m.n_nodes, n_own3 = n_owned >= 0 ? 3 * n_owned : nd;
    std::vector<double> xs(2 * (size_t)S.max_ncx), Xs(2 * (size_t)(x_all ? S.max_ncx : S.max_ncX)), stage(3 * (size_t)S.max_entries);
    for (int p = 0; p < S.n_patches; p++) {
        const uint8_t *b = &S.blob[(size_t)p * L.stride];
        const int32_t *hdr = reinterpret_cast<const int32_t *>(b);
        const int np = hdr[0], ncx = hdr[1] & 0xFFFF, ncX = hdr[1] >> 16, ne = hdr[2], nrows = hdr[3] & 0xFFFF;
        const uint32_t *cx = reinterpret_cast<const uint32_t *>(b + L.off_cx), *cX = reinterpret_cast<const uint32_t *>(b + L.off_cX);
        const uint32_t *et = reinterpret_cast<const uint32_t *>(b + L.off_et);
        const uint32_t *qn = reinterpret_cast<const uint32_t *>(b + L.off_qn);
        const uint8_t *ql = b + L.off_ql;
        const uint16_t *jo = reinterpret_cast<const uint16_t *>(b + L.off_jo);
        bool seen_ghost = false;
        std::fill(xs.begin(), xs.end(), 1e300);
        std::fill(Xs.begin(), Xs.end(), 1e300);
        // gather_chunks of matvec.cu: chunk c = doubles 2c, 2c+1; doubles beyond the owned range come from the landing
        // buffer (fused halo), doubles beyond the end of the vector are never copied
        for (int j = 0; j < ncx; j++) {
            if (j > 0 && cx[j] <= cx[j - 1]) { jfem_set_error("gather chunks of patch %d not ascending", p); throw 1; }
            for (int e = 0; e < 2; e++) {
                const long long d = 2LL * cx[j] + e;
                if (d >= nd) continue;
                const bool ghost = d >= n_own3;
                seen_ghost |= ghost;
                xs[2 * j + e] = ghost ? land[d - n_own3] : x[d];
                if (x_all) Xs[2 * j + e] = m.coords[d];
            }
        }
        if (!x_all)
            for (int j = 0; j < ncX; j++)
                for (int e = 0; e < 2; e++) {
                    const long long d = 2LL * cX[j] + e;
                    if (d < nd) Xs[2 * j + e] = m.coords[d];
                }
        // a chunk may drag in one double of a neighbouring (ghost) node that no element of the patch reads, so the flag may
        // only be set when a ghost double is present, and must be set when one is read: checked through the poison below
        if (((hdr[3] & 0x10000) != 0) && !seen_ghost) { jfem_set_error("ghost flag of patch %d set without ghost values", p); throw 1; }
        if (!(hdr[3] & 0x10000) && n_owned >= 0)   // not flagged: the kernel may run it before the halo has landed
            for (int j = 0; j < ncx; j++)
                for (int e = 0; e < 2; e++)
                    if (2LL * cx[j] + e >= n_own3 && 2LL * cx[j] + e < nd) xs[2 * j + e] = 1e300;
        std::fill(stage.begin(), stage.end(), 1e300);   // poison: every entry read must have been written
        for (int t = 0; t < ne; t++)
            element_phase<NNPE, CLS, OP_LINEAR, PtLinear, T>(pt, elem_offset + (long long)p * T + t, et, t, xs.data(), Xs.data(), nullptr, stage.data());
        (void)nrows;
        for (int q = 0; q < np; q++) {
            const uint32_t w = qn[q];
            const int len = ql[q];
            for (int c = 0; c < 3; c++) {
                double s = 0.0;
                for (int r = 0; r < len; r++) s += stage[3 * ((size_t)jo[r] + q) + c];
                if (project && ((w >> (PN_FIXSHIFT + c)) & 1u)) s = 0.0;
                if (w & PN_IFACE) ipart[3 * (size_t)(w & PN_ID_MASK) + c] = s;
                else y[3 * (size_t)(w & PN_ID_MASK) + c] = s;
            }
        }
    }
}

extern "C" const char *hostcheck_error() { return g_err; }

// y = K x through the patch tables, on the CPU.  conn 0-based.  stats[8]: wf_before, wf_after, wf_ideal, n_patches,
// n_interface_nodes, blob stride, n_partials, n_affine
extern "C" int hostcheck_matvec(int nnpe, long long n_nodes, long long n_elems, const double *coords, const int32_t *conn, const uint8_t *fixed, int EP,
                                int lane_window, int use_affine, double E, double nu, const double *x, double *y, int project, double *stats, long long n_owned) {
    MeshHost m;
    m.nnpe = nnpe; m.n_nodes = n_nodes; m.n_elems = n_elems;
    m.coords.assign(coords, coords + 3 * n_nodes);
    m.conn.assign(conn, conn + (size_t)nnpe * n_elems);
    m.fixed.assign(fixed, fixed + 3 * n_nodes);
    classify_elements(m, use_affine != 0);
    PatchSetHost sets[N_CLASSES];
    InterfaceHost hif;
    if (build_patch_sets(m, EP, use_affine != 0, lane_window, n_owned, sets, hif) != JFEM_OK) return 1;
    // embed the Dirichlet mask exactly like upload_fixed()
    for (int c = 0; c < N_CLASSES; c++) {
        PatchSetHost &S = sets[c];
        for (int p = 0; p < S.n_patches; p++) {
            uint32_t *qn = reinterpret_cast<uint32_t *>(&S.blob[(size_t)p * S.L.stride + S.L.off_qn]);
            for (int q = 0, nb = S.pnode_ptr[p]; q < S.pnode_ptr[p + 1] - nb; q++) {
                uint32_t w = S.qnodes[nb + q] & ~(7u << PN_FIXSHIFT);
                const long long id = S.qids[nb + q];
                for (int d = 0; d < 3; d++)
                    if (m.fixed[3 * id + d]) w |= (1u << (PN_FIXSHIFT + d));
                qn[q] = w;
            }
        }
    }
    PtLinear pt;
    pt.la = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu)); pt.mu = E / (2.0 * (1.0 + nu)); pt.sy = 0; pt.H = 0; pt.pe = nullptr; pt.pe_n = 0;
    for (long long i = 0; i < 3 * n_nodes; i++) y[i] = 1e300;   // poison: every dof must be written
    for (uint32_t n : hif.orphans) y[3 * (size_t)n] = y[3 * (size_t)n + 1] = y[3 * (size_t)n + 2] = 0.0;
    // partitioned replay: ghost entries of x are poisoned, their values live in the landing buffer only
    std::vector<double> xcopy(x, x + 3 * n_nodes), land(8, 0.0);
    if (n_owned >= 0 && n_owned < n_nodes) {
        land.assign(x + 3 * n_owned, x + 3 * n_nodes);
        for (long long i = 3 * n_owned; i < 3 * n_nodes; i++) xcopy[i] = 1e300;
    }
    const double *xin = xcopy.data();
    // ghost-reading patches must come after the others (the partial patch may stay last)
    for (int c = 0; c < N_CLASSES; c++) {
        const PatchSetHost &S = sets[c];
        const int nfull = (int)(S.n_elems / EP);
        for (int p = 1; p < nfull; p++)
            if (S.ghosty[p] < S.ghosty[p - 1]) { jfem_set_error("ghost-reading patches are not ordered last"); return 2; }
    }
    std::vector<double> ipart(3 * (size_t)hif.n_partials + 3, 1e300);
    long long off = 0;
    for (int c = 0; c < N_CLASSES; c++) {
        const PatchSetHost &S = sets[c];
        if (S.n_elems == 0) continue;
#define RUN(N, C, TT) run_set<N, C, TT>(S, off, m, pt, xin, y, project, ipart, n_owned, land.data())
#define RUN_T(N, C) do { if (EP == 128) RUN(N, C, 128); else if (EP == 256) RUN(N, C, 256); else RUN(N, C, 512); } while (0)
        try {
            if (nnpe == 10) { if (c == CLASS_AFFINE) RUN_T(10, CLASS_AFFINE); else RUN_T(10, CLASS_GENERAL); }
            else if (nnpe == 8) { if (c == CLASS_AFFINE) RUN_T(8, CLASS_AFFINE); else RUN_T(8, CLASS_GENERAL); }
            else RUN_T(4, CLASS_AFFINE);
        } catch (int) { return 3; }
        off += S.n_elems;
    }
    // interface nodes: slots in ascending (set, patch) order, like iface_reduce_kernel
    for (size_t i = 0; i < hif.inodes.size(); i++) {
        const long long b0 = hif.ibase[i], b1 = i + 1 < hif.inodes.size() ? hif.ibase[i + 1] : hif.n_partials;
        for (int c = 0; c < 3; c++) {
            double v = 0.0;
            for (long long r = b0; r < b1; r++) v += ipart[3 * (size_t)r + c];
            y[3 * (size_t)hif.inodes[i] + c] = v;
        }
    }
    const PatchSetHost &M = sets[CLASS_AFFINE].n_elems ? sets[CLASS_AFFINE] : sets[CLASS_GENERAL];
    stats[0] = M.wf_before; stats[1] = M.wf_after; stats[2] = M.wf_ideal;
    stats[3] = sets[0].n_patches + sets[1].n_patches; stats[4] = (double)hif.inodes.size(); stats[5] = M.L.stride;
    stats[6] = (double)hif.n_partials; stats[7] = (double)sets[CLASS_AFFINE].n_elems;
    return 0;
}

// wall time of the patch construction alone (setup cost study; not used by the tests)
#include <chrono>
extern "C" double hostcheck_build_seconds(int nnpe, long long n_nodes, long long n_elems, const double *coords, const int32_t *conn, int EP,
                                          int lane_window, long long n_owned, double *quality) {
    MeshHost m;
    m.nnpe = nnpe; m.n_nodes = n_nodes; m.n_elems = n_elems;
    m.coords.assign(coords, coords + 3 * n_nodes);
    m.conn.assign(conn, conn + (size_t)nnpe * n_elems);
    m.fixed.assign(3 * n_nodes, 0);
    auto t0 = std::chrono::steady_clock::now();
    classify_elements(m, true);
    PatchSetHost sets[N_CLASSES];
    InterfaceHost hif;
    if (build_patch_sets(m, EP, true, lane_window, n_owned, sets, hif) != JFEM_OK) return -1.0;
    if (quality) {   // patches, nodes per patch (mean, max), interface nodes, partial slots, max coordinate nodes, blob stride
        const PatchSetHost &S = sets[CLASS_AFFINE].n_elems ? sets[CLASS_AFFINE] : sets[CLASS_GENERAL];
        quality[0] = S.n_patches; quality[1] = S.n_patches ? (double)S.pnode_ptr[S.n_patches] / S.n_patches : 0; quality[2] = S.max_nodes;
        quality[3] = (double)hif.inodes.size(); quality[4] = (double)hif.n_partials; quality[5] = S.max_ncX; quality[6] = S.L.stride;
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---------------------------------------------------------------------------------------------------------------
// Replay of elem_warp_kernel (csrc/assemble.cu) for ONE element: "lanes" 0..NGP-1 run gp_geometry / gp_grad, then
// "lanes" 0..3*NNPE-1 build one column each with gp_column (csrc/asm_elem.cuh).  Ke is column-major (as the dense output
// of jfem_element_matrices).  kind: 0 LE, 1 Neo-Hookean, 2 J2 plasticity (st_old = 13 x NGP SoA), 3 St. Venant-Kirchhoff.
// ---------------------------------------------------------------------------------------------------------------
#include "../../juliafem.jl_b200/csrc/asm_elem.cuh"

struct HField {   // element-local field: 3 doubles per element node
    const double *base;
    JF_HD double operator()(int k, int c) const { return base[3 * k + c]; }
};

template <int NNPE, class Pt>
static bool columns_one(const Pt &pt0, const double *X, const double *u, double *Ke) {
    constexpr int ND = 3 * NNPE, NGP = ElemRule<NNPE>::NGP;
    double gN[NGP][3 * NNPE], Gu[NGP][9], w[NGP];
    HField XF{X}, UF{u};
    for (int g = 0; g < NGP; g++) {
        w[g] = gp_geometry<NNPE>(g, XF, gN[g]);
        if (Pt::NF == 2) gp_grad<NNPE>(UF, gN[g], Gu[g]);
    }
    bool ok = true;
    for (int lane = 0; lane < ND; lane++) {
        const int l = lane / 3, cj = lane - 3 * l;
        Pt pt = pt0;
        pt.load(0);
        double acc[NNPE][3];
        for (int k = 0; k < NNPE; k++) acc[k][0] = acc[k][1] = acc[k][2] = 0.0;
        for (int g = 0; g < NGP; g++) ok &= gp_column<NNPE>(pt, (long long)g, gN[g], w[g], Gu[g], l, cj, acc);
        for (int k = 0; k < NNPE; k++) for (int c = 0; c < 3; c++) Ke[(size_t)lane * ND + 3 * k + c] = acc[k][c];
    }
    return ok;
}

template <int NNPE>
static int columns_kind(int kind, const double *par, int geo, const double *st_old, const double *X, const double *u, double *Ke) {
    MatBase mb;
    mb.la = par[0] * par[1] / ((1.0 + par[1]) * (1.0 - 2.0 * par[1])); mb.mu = par[0] / (2.0 * (1.0 + par[1]));
    mb.sy = par[2]; mb.H = par[3]; mb.pe = nullptr; mb.pe_n = 0;
    bool ok;
    if (kind == 0) { PtLinear pt; static_cast<MatBase &>(pt) = mb; ok = columns_one<NNPE>(pt, X, u, Ke); }
    else if (kind == 1) { PtNHTangent pt; static_cast<MatBase &>(pt) = mb; ok = columns_one<NNPE>(pt, X, u, Ke); }
    else if (kind == 2) { PtPPTangent pt; static_cast<MatBase &>(pt) = mb; pt.st_old = st_old; pt.n_gp = ElemRule<NNPE>::NGP; ok = columns_one<NNPE>(pt, X, u, Ke); }
    else { PtStVKTangent pt; static_cast<MatBase &>(pt) = mb; pt.geo = geo; ok = columns_one<NNPE>(pt, X, u, Ke); }
    return ok ? 0 : 1;
}

extern "C" int hostcheck_element_columns(int nnpe, const double *X, const double *u, int kind, const double *par, int geo, const double *st_old,
                                         double *Ke) {
    if (nnpe == 10) return columns_kind<10>(kind, par, geo, st_old, X, u, Ke);
    if (nnpe == 8) return columns_kind<8>(kind, par, geo, st_old, X, u, Ke);
    return columns_kind<4>(kind, par, geo, st_old, X, u, Ke);
}

// ---------------------------------------------------------------------------------------------------------------
// Host side of the assembled path (patches.cpp: build_node_adjacency, greedy_colouring), expanded to the reference's CSR
// pattern exactly like expand_pattern_kernel of csrc/assemble.cu.  rowptr: 3 n_nodes + 1; colind (may be null): nnz;
// colour (may be null): colour of every element; times[0..1] = seconds of adjacency / colouring.
// Returns nnz, or -1.  Also verifies eblk (every (k, l) entry points at node l inside the row of node k).
// ---------------------------------------------------------------------------------------------------------------
extern "C" long long hostcheck_pattern(int nnpe, long long n_nodes, long long n_elems, const int32_t *conn, long long *rowptr, int32_t *colind,
                                       int32_t *colour, double *times) {
    MeshHost m;
    m.nnpe = nnpe; m.n_nodes = n_nodes; m.n_elems = n_elems;
    m.conn.assign(conn, conn + (size_t)nnpe * n_elems);
    std::vector<int64_t> ap, cptr;
    std::vector<int32_t> adj, celems;
    std::vector<uint16_t> eblk;
    auto t0 = std::chrono::steady_clock::now();
    if (build_node_adjacency(m, ap, adj, eblk) != JFEM_OK) return -1;
    auto t1 = std::chrono::steady_clock::now();
    greedy_colouring(m, cptr, celems);
    auto t2 = std::chrono::steady_clock::now();
    if (times) { times[0] = std::chrono::duration<double>(t1 - t0).count(); times[1] = std::chrono::duration<double>(t2 - t1).count(); }
    for (long long e = 0; e < n_elems; e++)
        for (int k = 0; k < nnpe; k++)
            for (int l = 0; l < nnpe; l++)
                if (adj[ap[conn[e * nnpe + k]] + eblk[(e * nnpe + k) * nnpe + l]] != conn[e * nnpe + l]) { jfem_set_error("eblk entry wrong at element %lld", e); return -1; }
    for (long long a = 0; a < n_nodes; a++) {
        const long long p = ap[a], deg = ap[a + 1] - p;
        for (int c = 0; c < 3; c++) {
            const long long r0 = 9 * p + 3 * c * deg;
            rowptr[3 * a + c] = r0;
            if (colind)
                for (long long q = 0; q < deg; q++)
                    for (int d = 0; d < 3; d++) colind[r0 + 3 * q + d] = 3 * adj[p + q] + d;
        }
    }
    rowptr[3 * n_nodes] = 9 * ap[n_nodes];
    if (colour)
        for (size_t c = 0; c + 1 < cptr.size(); c++)
            for (long long q = cptr[c]; q < cptr[c + 1]; q++) colour[celems[q]] = (int32_t)c;
    return 9 * ap[n_nodes];
}

// loads.cu -- the steps right before and right after the solve:
//   * consistent external loads: body load (src/problems_elasticity.jl:412-426), surface traction / pressure on Tri3, Tri6,
//     Quad4 faces (src/problems_elasticity.jl:454-502).  The reference's GPU precedent (apply_surface_traction_kernel!,
//     ext/JuliaFEMCUDAExt.jl:368-416) lumps Tri3 areas with atomics; here the integration is the consistent one of the CPU
//     path and every node GATHERS from its incident elements / faces in ascending order: deterministic, no atomics.
//   * reaction forces on the constrained dofs (src/solvers.jl:205-216)
//   * nodal least-squares recovery of Gauss-point strain / stress (lsq_fit, src/problems_elasticity.jl:547-594)
//   * penalty Dirichlet conditions on the assembled CSR (src/element_assembly_structures.jl:237-252)
// Shape functions are written in the reference's expanded-monomial form (src/basis/lagrange_generated.jl).
#include <algorithm>
#include <vector>

#include "elem.cuh"
#include "handle.h"

using namespace jf;

void fill_material(const jfem_handle *h, jf::MatBase &m);

namespace {

// ---------------------------------------------------------------- shape functions / rules (host + device)
// volume: Tet4 (lagrange_generated.jl:239-254), Tet10 (:267-282), Hex8 (:295-310); N[i], dN[i][a] = dN_i/dxi_a
JF_HD void vol_shape(int nn, const double (&xi)[3], double *N, double (*dN)[3]) {
    const double u = xi[0], v = xi[1], w = xi[2];
    if (nn == 4) {
        N[0] = 1 - u - v - w; N[1] = u; N[2] = v; N[3] = w;
        const double d[4][3] = {{-1, -1, -1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int i = 0; i < 4; i++) for (int a = 0; a < 3; a++) dN[i][a] = d[i][a];
    } else if (nn == 10) {
        N[0] = 1 - 3 * u - 3 * v - 3 * w + 2 * u * u + 2 * v * v + 2 * w * w + 4 * u * v + 4 * u * w + 4 * v * w;
        N[1] = -u + 2 * u * u; N[2] = -v + 2 * v * v; N[3] = -w + 2 * w * w;
        N[4] = 4 * u - 4 * u * u - 4 * u * v - 4 * u * w; N[5] = 4 * u * v; N[6] = 4 * v - 4 * v * v - 4 * u * v - 4 * v * w;
        N[7] = 4 * w - 4 * w * w - 4 * u * w - 4 * v * w; N[8] = 4 * u * w; N[9] = 4 * v * w;
        const double d0 = -3 + 4 * u + 4 * v + 4 * w;
        const double d[10][3] = {{d0, d0, d0}, {-1 + 4 * u, 0, 0}, {0, -1 + 4 * v, 0}, {0, 0, -1 + 4 * w},
                                 {4 - 8 * u - 4 * v - 4 * w, -4 * u, -4 * u}, {4 * v, 4 * u, 0}, {-4 * v, 4 - 8 * v - 4 * u - 4 * w, -4 * v},
                                 {-4 * w, -4 * w, 4 - 8 * w - 4 * u - 4 * v}, {4 * w, 0, 4 * u}, {0, 4 * w, 4 * v}};
        for (int i = 0; i < 10; i++) for (int a = 0; a < 3; a++) dN[i][a] = d[i][a];
    } else {
        const double s[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
        for (int i = 0; i < 8; i++) {
            const double a = s[i][0], b = s[i][1], c = s[i][2];
            N[i] = 0.125 * (1 + a * u) * (1 + b * v) * (1 + c * w);
            dN[i][0] = 0.125 * a * (1 + b * v) * (1 + c * w);
            dN[i][1] = 0.125 * b * (1 + a * u) * (1 + c * w);
            dN[i][2] = 0.125 * c * (1 + a * u) * (1 + b * v);
        }
    }
}
// default rules (src/elements/integrate.jl:19,26-27): Tet4 GLTET1, Tet10 GLTET4, Hex8 GLHEX8 (first index fastest)
JF_HD int vol_rule(int nn, double *w, double (*xi)[3]) {
    if (nn == 4) { w[0] = 1.0 / 6.0; xi[0][0] = xi[0][1] = xi[0][2] = 0.25; return 1; }
    if (nn == 10) {
        for (int g = 0; g < 4; g++) {
            w[g] = 1.0 / 24.0;
            for (int a = 0; a < 3; a++) xi[g][a] = (a == g) ? T10_A : T10_B;     // (a,b,b) (b,a,b) (b,b,a) (b,b,b), gltet.jl:18-25
        }
        return 4;
    }
    for (int g = 0; g < 8; g++) {
        w[g] = 1.0;
        xi[g][0] = (g & 1) ? HEX_GA : -HEX_GA; xi[g][1] = (g & 2) ? HEX_GA : -HEX_GA; xi[g][2] = (g & 4) ? HEX_GA : -HEX_GA;
    }
    return 8;
}
// Rule for the least-squares recovery.  The reference integrates lsq_fit with the element's DEFAULT rule (pe:555), which
// makes the mass matrix sum w N N' of Tet4 / Tet10 exactly singular (GLTET1 / GLTET4: 1 / 4 points for 4 / 10 functions per
// element; on the reference's own tet10.inp and JuliaFEMSMP18 meshes it has zero eigenvalues) -- a defect, not followed.
// Used instead: the next rule of the reference's own integration_rule_mapping that integrates N N' exactly
// (src/elements/integrate.jl:5-10,29-30: "sometimes we want to increase integration order, e.g. when integrating mass
// matrix"): Tet4 -> GLTET4, Tet10 -> GLTET15 (src/quadrature/gltet.jl:44-64); Hex8 keeps GLHEX8 (regular there).
JF_HD int mass_rule(int nn, double *w, double (*xi)[3]) {
    if (nn == 8) return vol_rule(8, w, xi);
    if (nn == 4) return vol_rule(10, w, xi);
    const double s15 = 3.872983346207417;   // sqrt(15)
    const double a = 0.25, b1 = (7.0 + s15) / 34.0, b2 = (7.0 - s15) / 34.0, c1 = (13.0 - 3.0 * s15) / 34.0, c2 = (13.0 + 3.0 * s15) / 34.0;
    const double d = (5.0 - s15) / 20.0, f = (5.0 + s15) / 20.0;
    const double w1 = 8.0 / 405.0, w2 = (2665.0 - 14.0 * s15) / 226800.0, w3 = (2665.0 + 14.0 * s15) / 226800.0, w4 = 5.0 / 567.0;
    const double p[15][3] = {{a, a, a}, {b1, b1, b1}, {b1, b1, c1}, {b1, c1, b1}, {c1, b1, b1}, {b2, b2, b2}, {b2, b2, c2}, {b2, c2, b2},
                             {c2, b2, b2}, {d, d, f}, {d, f, d}, {f, d, d}, {d, f, f}, {f, d, f}, {f, f, d}};
    const double ww[15] = {w1, w2, w2, w2, w2, w3, w3, w3, w3, w4, w4, w4, w4, w4, w4};
    for (int g = 0; g < 15; g++) { w[g] = ww[g]; for (int c = 0; c < 3; c++) xi[g][c] = p[g][c]; }
    return 15;
}
// surface: Tri3 (:99-115), Tri6 (:127-143), Quad4 (:155-171); rules GLTRI1 / GLTRI3 (src/quadrature/gltri.jl:7-27) / GLQUAD4
JF_HD void surf_shape(int nn, double u, double v, double *N, double (*dN)[2]) {
    if (nn == 3) {
        N[0] = 1 - u - v; N[1] = u; N[2] = v;
        dN[0][0] = -1; dN[0][1] = -1; dN[1][0] = 1; dN[1][1] = 0; dN[2][0] = 0; dN[2][1] = 1;
    } else if (nn == 6) {
        N[0] = 1 - 3 * u - 3 * v + 2 * u * u + 4 * u * v + 2 * v * v; N[1] = -u + 2 * u * u; N[2] = -v + 2 * v * v;
        N[3] = 4 * u - 4 * u * u - 4 * u * v; N[4] = 4 * u * v; N[5] = 4 * v - 4 * u * v - 4 * v * v;
        dN[0][0] = -3 + 4 * u + 4 * v; dN[0][1] = -3 + 4 * u + 4 * v;
        dN[1][0] = -1 + 4 * u; dN[1][1] = 0; dN[2][0] = 0; dN[2][1] = -1 + 4 * v;
        dN[3][0] = 4 - 8 * u - 4 * v; dN[3][1] = -4 * u; dN[4][0] = 4 * v; dN[4][1] = 4 * u;
        dN[5][0] = -4 * v; dN[5][1] = 4 - 4 * u - 8 * v;
    } else {
        const double s[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};
        for (int i = 0; i < 4; i++) {
            N[i] = 0.25 * (1 + s[i][0] * u) * (1 + s[i][1] * v);
            dN[i][0] = 0.25 * s[i][0] * (1 + s[i][1] * v); dN[i][1] = 0.25 * s[i][1] * (1 + s[i][0] * u);
        }
    }
}
JF_HD int surf_rule(int nn, double *w, double (*xi)[2]) {
    if (nn == 3) { w[0] = 0.5; xi[0][0] = xi[0][1] = 1.0 / 3.0; return 1; }
    if (nn == 6) {
        const double p[3][2] = {{2.0 / 3.0, 1.0 / 6.0}, {1.0 / 6.0, 2.0 / 3.0}, {1.0 / 6.0, 1.0 / 6.0}};
        for (int g = 0; g < 3; g++) { w[g] = 1.0 / 6.0; xi[g][0] = p[g][0]; xi[g][1] = p[g][1]; }
        return 3;
    }
    for (int g = 0; g < 4; g++) { w[g] = 1.0; xi[g][0] = (g & 1) ? HEX_GA : -HEX_GA; xi[g][1] = (g & 2) ? HEX_GA : -HEX_GA; }
    return 4;
}

// J[a][b] = sum_i dN_i[a] X_i[b] (src/basis/math.jl:47-54)
template <int MAXN>
__device__ __forceinline__ double vol_jacobian(int nn, const double (*dN)[3], const double (*X)[3], double (&J)[3][3]) {
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
        double s = 0;
        for (int i = 0; i < nn; i++) s += dN[i][a] * X[i][b];
        J[a][b] = s;
    }
    return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
           J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
}

// ---------------------------------------------------------------- body load: one thread per node
// inc: node -> (element * nn + local node) incidences, ascending.  b: 3 values or 3 per element.
__global__ void body_load_kernel(int nn, long long n_nodes, const long long *__restrict__ nptr, const long long *__restrict__ inc,
                                 const int32_t *__restrict__ conn, const double *__restrict__ coords, const double *__restrict__ b, int per_elem,
                                 int accumulate, double *__restrict__ f) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_nodes) return;
    double acc[3] = {0, 0, 0};
    double w[8], xi[8][3];
    const int ng = vol_rule(nn, w, xi);
    for (long long p = nptr[n]; p < nptr[n + 1]; p++) {
        const long long e = inc[p] / nn;
        const int k = (int)(inc[p] - e * nn);
        double X[10][3];
        for (int i = 0; i < nn; i++) for (int c = 0; c < 3; c++) X[i][c] = coords[3LL * conn[e * nn + i] + c];
        const double *be = per_elem ? b + 3 * e : b;
        for (int g = 0; g < ng; g++) {
            double N[10], dN[10][3], J[3][3];
            vol_shape(nn, xi[g], N, dN);
            const double wd = w[g] * vol_jacobian<10>(nn, dN, X, J);
            for (int c = 0; c < 3; c++) acc[c] += wd * N[k] * be[c];
        }
    }
    for (int c = 0; c < 3; c++) f[3 * n + c] = (accumulate ? f[3 * n + c] : 0.0) + acc[c];
}

// ---------------------------------------------------------------- surface loads: one thread per node of the face set
// finc: node slot -> (face * nn + local node) incidences; fnode[slot] = node id
__global__ void surface_load_kernel(int nn, long long n_slots, const int32_t *__restrict__ fnode, const long long *__restrict__ fptr,
                                    const long long *__restrict__ finc, const int32_t *__restrict__ faces, const double *__restrict__ coords,
                                    const double *__restrict__ traction, const double *__restrict__ pressure, double *__restrict__ f) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slots) return;
    double acc[3] = {0, 0, 0};
    double w[4], xi[4][2];
    const int ng = surf_rule(nn, w, xi);
    for (long long p = fptr[s]; p < fptr[s + 1]; p++) {
        const long long fc = finc[p] / nn;
        const int k = (int)(finc[p] - fc * nn);
        double X[6][3];
        for (int i = 0; i < nn; i++) for (int c = 0; c < 3; c++) X[i][c] = coords[3LL * faces[fc * nn + i] + c];
        for (int g = 0; g < ng; g++) {
            double N[6], dN[6][2], t1[3] = {0, 0, 0}, t2[3] = {0, 0, 0};
            surf_shape(nn, xi[g][0], xi[g][1], N, dN);
            for (int i = 0; i < nn; i++) for (int c = 0; c < 3; c++) { t1[c] += dN[i][0] * X[i][c]; t2[c] += dN[i][1] * X[i][c]; }
            const double nx = t1[1] * t2[2] - t1[2] * t2[1], ny = t1[2] * t2[0] - t1[0] * t2[2], nz = t1[0] * t2[1] - t1[1] * t2[0];
            const double detJ = sqrt(nx * nx + ny * ny + nz * nz);     // || dX/dxi1 x dX/dxi2 ||  (src/elements/elements.jl:799-810)
            const double wd = w[g] * detJ;
            if (traction) for (int c = 0; c < 3; c++) acc[c] += wd * N[k] * traction[3 * fc + c];          // pe:472-475
            if (pressure) {                                                                                   // pe:485-491
                const double q = -pressure[fc] * wd * N[k] / detJ;
                acc[0] += q * nx; acc[1] += q * ny; acc[2] += q * nz;
            }
        }
    }
    const long long n = fnode[s];
    for (int c = 0; c < 3; c++) f[3 * n + c] += acc[c];
}

__global__ void reactions_kernel(long long n, const uint8_t *__restrict__ fixed, const double *__restrict__ fint, const double *__restrict__ fext,
                                 double *__restrict__ la) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) la[i] = fixed[i] ? fint[i] - (fext ? fext[i] : 0.0) : 0.0;
}

// ---------------------------------------------------------------- least-squares recovery (lsq_fit, pe:547-594)
// Mass matrix on the node adjacency pattern, coloured element launches (race-free, ordered): M[k,l] += w detJ N_k N_l
__global__ void mass_assemble_kernel(int nn, long long ne, const int32_t *__restrict__ elems, long long e0, const int32_t *__restrict__ conn,
                                     const double *__restrict__ coords, const long long *__restrict__ adjptr, const uint16_t *__restrict__ eblk,
                                     double *__restrict__ M) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne) return;
    const long long e = elems[e0 + i];
    double X[10][3], w[15], xi[15][3];
    int nd[10];
    for (int k = 0; k < nn; k++) { nd[k] = conn[e * nn + k]; for (int c = 0; c < 3; c++) X[k][c] = coords[3LL * nd[k] + c]; }
    const int ng = mass_rule(nn, w, xi);
    for (int g = 0; g < ng; g++) {
        double N[10], dN[10][3], J[3][3];
        vol_shape(nn, xi[g], N, dN);
        const double wd = w[g] * vol_jacobian<10>(nn, dN, X, J);
        for (int k = 0; k < nn; k++) {
            double *row = M + adjptr[nd[k]];
            for (int l = 0; l < nn; l++) row[eblk[(e * nn + k) * nn + l]] += wd * N[k] * N[l];
        }
    }
}
// right-hand sides b[node][i] += w detJ f_i N  with f = strain or linear-elastic stress vector (pe:520-545); material per element
__global__ void recover_rhs_kernel(int nn, long long ne, const int32_t *__restrict__ elems, long long e0, const int32_t *__restrict__ conn,
                                   const double *__restrict__ coords, const double *__restrict__ u, MatBase mat, const long long *__restrict__ e2i,
                                   int field, double *__restrict__ b) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ne) return;
    const long long e = elems[e0 + i];
    double X[10][3], U[10][3], w[15], xi[15][3];
    int nd[10];
    for (int k = 0; k < nn; k++) {
        nd[k] = conn[e * nn + k];
        for (int c = 0; c < 3; c++) { X[k][c] = coords[3LL * nd[k] + c]; U[k][c] = u[3LL * nd[k] + c]; }
    }
    mat.load(e2i[e]);
    const int ng = mass_rule(nn, w, xi);
    for (int g = 0; g < ng; g++) {
        double N[10], dN[10][3], J[3][3], iJ[3][3];
        vol_shape(nn, xi[g], N, dN);
        const double wd = w[g] * vol_jacobian<10>(nn, dN, X, J);
        inv3x3(J, iJ);
        double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};     // grad u = sum u_k (x) (inv(J) dN_k)   (math.jl:198-201,249-255)
        for (int k = 0; k < nn; k++) {
            double gk[3];
            for (int bb = 0; bb < 3; bb++) gk[bb] = iJ[bb][0] * dN[k][0] + iJ[bb][1] * dN[k][1] + iJ[bb][2] * dN[k][2];
            for (int a = 0; a < 3; a++) for (int bb = 0; bb < 3; bb++) G[a][bb] += U[k][a] * gk[bb];
        }
        double f[6] = {G[0][0], G[1][1], G[2][2], 0.5 * (G[0][1] + G[1][0]), 0.5 * (G[1][2] + G[2][1]), 0.5 * (G[0][2] + G[2][0])};
        if (field == JFEM_FIELD_STRESS) {
            const double tr = mat.la * (f[0] + f[1] + f[2]);
            for (int q = 0; q < 3; q++) f[q] = tr + 2 * mat.mu * f[q];
            for (int q = 3; q < 6; q++) f[q] = 2 * mat.mu * f[q];
        }
        for (int k = 0; k < nn; k++) for (int q = 0; q < 6; q++) b[6LL * nd[k] + q] += wd * N[k] * f[q];
    }
}

// ---- Jacobi-PCG on the scalar node-adjacency CSR with 6 right-hand sides at once (vectors are [node][6])
__global__ void mass_diag_kernel(long long n, const long long *__restrict__ adjptr, const int32_t *__restrict__ adj, const double *__restrict__ M,
                                 double *__restrict__ dinv) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double d = 0;
    for (long long p = adjptr[r]; p < adjptr[r + 1]; p++) if (adj[p] == r) d = M[p];
    dinv[r] = d != 0 ? 1.0 / d : 0.0;     // node without elements: the reference drops such rows (get_nonzero_rows, pe:567)
}
__global__ void mass_spmv6_kernel(long long n, const long long *__restrict__ adjptr, const int32_t *__restrict__ adj, const double *__restrict__ M,
                                  const double *__restrict__ x, double *__restrict__ y) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (long long p = adjptr[r]; p < adjptr[r + 1]; p++) {
        const double m = M[p];
        const double *xc = x + 6LL * adj[p];
        for (int q = 0; q < 6; q++) s[q] += m * xc[q];
    }
    for (int q = 0; q < 6; q++) y[6 * r + q] = s[q];
}
// out[q] = sum_r a[r][q] * b[r][q] (* w[r] if w): one block, fixed order -> deterministic (vectors are small: n_nodes)
__global__ void dot6_kernel(long long n, const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ wgt, double *__restrict__ out) {
    __shared__ double sh[256][6];
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (long long r = threadIdx.x; r < n; r += blockDim.x)
        for (int q = 0; q < 6; q++) s[q] += a[6 * r + q] * b[6 * r + q] * (wgt ? wgt[r] : 1.0);
    for (int q = 0; q < 6; q++) sh[threadIdx.x][q] = s[q];
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) for (int q = 0; q < 6; q++) sh[threadIdx.x][q] += sh[threadIdx.x + o][q];
        __syncthreads();
    }
    if (threadIdx.x < 6) out[threadIdx.x] = sh[0][threadIdx.x];
}
// x += alpha p ; r -= alpha Ap  (alpha per right-hand side)
__global__ void pcg6_update_kernel(long long n, const double *__restrict__ alpha, const double *__restrict__ p, const double *__restrict__ Ap,
                                   double *__restrict__ x, double *__restrict__ r) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6 * n) return;
    const double a = alpha[i % 6];
    x[i] += a * p[i];
    r[i] -= a * Ap[i];
}
// p = dinv r + beta p
__global__ void pcg6_dir_kernel(long long n, const double *__restrict__ beta, const double *__restrict__ dinv, const double *__restrict__ r,
                                double *__restrict__ p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 6 * n) return;
    p[i] = dinv[i / 6] * r[i] + beta[i % 6] * p[i];
}

// ---------------------------------------------------------------- penalty Dirichlet on the CSR (eas:237-252)
__global__ void absmax_kernel(long long n, const double *__restrict__ v, double *__restrict__ out) {   // one block
    __shared__ double sh[256];
    double m = 0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, fabs(v[i]));
    sh[threadIdx.x] = m;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}
__global__ void penalty_kernel(long long n_rows, const long long *__restrict__ rowptr, const int32_t *__restrict__ colind, const uint8_t *__restrict__ fixed,
                               const double *__restrict__ prescribed, const double *__restrict__ maxabs, double scale, double *__restrict__ vals,
                               double *__restrict__ rhs) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows || !fixed[r]) return;
    const double pen = scale * *maxabs;
    for (long long p = rowptr[r]; p < rowptr[r + 1]; p++) if (colind[p] == r) vals[p] += pen;
    if (rhs) rhs[r] = pen * prescribed[r];
}

// node -> incidences (item * nn + local), ascending item order
void incidences(int nn, long long n_nodes, long long n_items, const int32_t *conn, std::vector<long long> &ptr, std::vector<long long> &inc) {
    ptr.assign(n_nodes + 1, 0);
    for (long long i = 0; i < n_items * nn; i++) ptr[conn[i] + 1]++;
    for (long long i = 0; i < n_nodes; i++) ptr[i + 1] += ptr[i];
    inc.resize(ptr[n_nodes]);
    std::vector<long long> fill(ptr.begin(), ptr.end() - 1);
    for (long long e = 0; e < n_items; e++) for (int k = 0; k < nn; k++) inc[fill[conn[e * nn + k]]++] = e * nn + k;
}

int ensure_incidences(jfem_handle *h) {
    if (h->n2e_ptr.n) return JFEM_OK;
    std::vector<long long> ptr, inc;
    incidences(h->mesh.nnpe, h->mesh.n_nodes, h->mesh.n_elems, h->mesh.conn.data(), ptr, inc);
    JFEM_TRY(h->n2e_ptr.upload(ptr));
    JFEM_TRY(h->n2e_inc.upload(inc));
    if (h->dconn.n == 0) JFEM_TRY(h->dconn.upload(h->mesh.conn));
    return JFEM_OK;
}

// stage helpers for host / device vectors of arbitrary length
int to_dev(jfem_handle *h, const double *p, size_t n, int on_device, DevBuf<double> &buf, const double **out) {
    if (!p) { *out = nullptr; return JFEM_OK; }
    if (on_device) { *out = p; return JFEM_OK; }
    if (buf.n < n) JFEM_TRY(buf.alloc(n));
    JFEM_CUDA(cudaMemcpyAsync(buf.p, p, n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    *out = buf.p;
    return JFEM_OK;
}

}  // namespace

#define CHECK_H(h)                                             \
    do {                                                       \
        if (!(h)) { jfem_set_error("null handle"); return JFEM_EINVAL; } \
        JFEM_CUDA(cudaSetDevice((h)->device));                 \
    } while (0)

extern "C" {

int jfem_body_load(jfem_handle *h, const double *b, int per_element, int accumulate, double *f, int on_device) {
    CHECK_H(h);
    if (!b || !f) { jfem_set_error("jfem_body_load: null argument"); return JFEM_EINVAL; }
    JFEM_TRY(ensure_built(h));
    JFEM_TRY(ensure_incidences(h));
    const long long nn = h->mesh.n_nodes;
    const size_t nd = (size_t)h->n_dofs();
    DevBuf<double> db;
    JFEM_TRY(db.upload(std::vector<double>(b, b + (per_element ? 3 * (size_t)h->mesh.n_elems : 3))));
    double *df = f;
    if (!on_device) {
        if (h->wy.n != nd) JFEM_TRY(h->wy.alloc(nd));
        df = h->wy.p;
        if (accumulate) JFEM_CUDA(cudaMemcpyAsync(df, f, nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    body_load_kernel<<<(unsigned)((nn + 127) / 128), 128, 0, h->stream>>>(h->mesh.nnpe, nn, h->n2e_ptr.p, h->n2e_inc.p, h->dconn.p, h->coords.p, db.p,
                                                                          per_element ? 1 : 0, accumulate ? 1 : 0, df);
    JFEM_CUDA(cudaGetLastError());
    h->total_launches++;
    if (!on_device) JFEM_CUDA(cudaMemcpyAsync(f, df, nd * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));   // db is released on return
    db.release();
    return JFEM_OK;
}

int jfem_surface_load(jfem_handle *h, int face_type, int64_t n_faces, const int32_t *faces, const double *traction, const double *pressure,
                      int accumulate, double *f, int on_device) {
    CHECK_H(h);
    if (face_type != JFEM_TRI3 && face_type != JFEM_TRI6 && face_type != JFEM_QUAD4) {
        jfem_set_error("unsupported surface element type %d (supported: Tri3=3, Quad4=4, Tri6=6)", face_type); return JFEM_EINVAL;
    }
    if (n_faces < 0 || (n_faces && !faces) || !f) { jfem_set_error("jfem_surface_load: bad arguments"); return JFEM_EINVAL; }
    JFEM_TRY(ensure_built(h));
    const int nn = face_type;
    const size_t nd = (size_t)h->n_dofs();
    std::vector<int32_t> fc((size_t)n_faces * nn);
    for (size_t i = 0; i < fc.size(); i++) {
        const long long v = (long long)faces[i] - h->index_base;
        if (v < 0 || v >= h->mesh.n_nodes) { jfem_set_error("face node %lld out of range", (long long)faces[i]); return JFEM_EINVAL; }
        fc[i] = (int32_t)v;
    }
    // compact node slots of the face set: slot -> node, slot -> incidences (ascending face order)
    std::vector<int32_t> nodes(fc);
    std::sort(nodes.begin(), nodes.end());
    nodes.erase(std::unique(nodes.begin(), nodes.end()), nodes.end());
    std::vector<int32_t> slot_conn(fc.size());
    for (size_t i = 0; i < fc.size(); i++) slot_conn[i] = (int32_t)(std::lower_bound(nodes.begin(), nodes.end(), fc[i]) - nodes.begin());
    std::vector<long long> ptr, inc;
    incidences(nn, (long long)nodes.size(), n_faces, slot_conn.data(), ptr, inc);
    DevBuf<int32_t> dnodes, dfaces;
    DevBuf<long long> dptr, dinc;
    DevBuf<double> dt, dp;
    JFEM_TRY(dnodes.upload(nodes)); JFEM_TRY(dfaces.upload(fc)); JFEM_TRY(dptr.upload(ptr)); JFEM_TRY(dinc.upload(inc));
    if (traction) JFEM_TRY(dt.upload(std::vector<double>(traction, traction + 3 * (size_t)n_faces)));
    if (pressure) JFEM_TRY(dp.upload(std::vector<double>(pressure, pressure + (size_t)n_faces)));
    double *df = f;
    if (!on_device) {
        if (h->wy.n != nd) JFEM_TRY(h->wy.alloc(nd));
        df = h->wy.p;
        if (accumulate) JFEM_CUDA(cudaMemcpyAsync(df, f, nd * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    if (!accumulate) JFEM_CUDA(cudaMemsetAsync(df, 0, nd * sizeof(double), h->stream));
    const long long ns = (long long)nodes.size();
    if (ns) {
        surface_load_kernel<<<(unsigned)((ns + 127) / 128), 128, 0, h->stream>>>(nn, ns, dnodes.p, dptr.p, dinc.p, dfaces.p, h->coords.p, dt.p, dp.p, df);
        JFEM_CUDA(cudaGetLastError());
        h->total_launches++;
    }
    if (!on_device) JFEM_CUDA(cudaMemcpyAsync(f, df, nd * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    dnodes.release(); dfaces.release(); dptr.release(); dinc.release(); dt.release(); dp.release();
    return JFEM_OK;
}

int jfem_reactions(jfem_handle *h, const double *u, const double *f_ext, double *la, int on_device) {
    CHECK_H(h);
    if (!u || !la) { jfem_set_error("jfem_reactions: null vector"); return JFEM_EINVAL; }
    JFEM_TRY(ensure_built(h));
    const size_t nd = (size_t)h->n_dofs();
    const double *du, *dfx;
    JFEM_TRY(to_dev(h, u, nd, on_device, h->wx, &du));
    DevBuf<double> fx, fint, dla;
    JFEM_TRY(to_dev(h, f_ext, nd, on_device, fx, &dfx));
    JFEM_TRY(fint.alloc(nd));
    if (h->n_ranks > 1) JFEM_TRY(halo_exchange(h, (double *)du, true));
    JFEM_TRY(op_apply(h, OP_RESIDUAL, du, fint.p, 0, nullptr));     // f_int(u) (== K u for the linear operator), unprojected
    double *out = la;
    if (!on_device) { JFEM_TRY(dla.alloc(nd)); out = dla.p; }
    reactions_kernel<<<(unsigned)((nd + 255) / 256), 256, 0, h->stream>>>((long long)nd, h->fixed.p, fint.p, dfx, out);
    JFEM_CUDA(cudaGetLastError());
    h->total_launches++;
    if (!on_device) JFEM_CUDA(cudaMemcpyAsync(la, out, nd * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    fx.release(); fint.release(); dla.release();
    return JFEM_OK;
}

int jfem_nodal_recover(jfem_handle *h, const double *u, int field, double *out, int on_device) {
    CHECK_H(h);
    if (!u || !out || (field != JFEM_FIELD_STRAIN && field != JFEM_FIELD_STRESS)) { jfem_set_error("jfem_nodal_recover: bad arguments"); return JFEM_EINVAL; }
    if (h->mat_kind < 0) { jfem_set_error("jfem_set_material has not been called"); return JFEM_ESTATE; }
    JFEM_TRY(csr_build(h));          // node adjacency pattern, eblk, colouring, device connectivity
    const long long nn = h->mesh.n_nodes;
    const int nnpe = h->mesh.nnpe;
    const size_t nd = (size_t)h->n_dofs(), nadj = (size_t)h->h_nadj_ptr[nn];
    const double *du;
    JFEM_TRY(to_dev(h, u, nd, on_device, h->wx, &du));
    DevBuf<double> M, b, x, r, p, Ap, dinv, sc;
    JFEM_TRY(M.alloc(nadj)); JFEM_TRY(b.alloc(6 * (size_t)nn)); JFEM_TRY(x.alloc(6 * (size_t)nn)); JFEM_TRY(r.alloc(6 * (size_t)nn));
    JFEM_TRY(p.alloc(6 * (size_t)nn)); JFEM_TRY(Ap.alloc(6 * (size_t)nn)); JFEM_TRY(dinv.alloc((size_t)nn)); JFEM_TRY(sc.alloc(24));
    JFEM_CUDA(cudaMemsetAsync(M.p, 0, M.bytes(), h->stream));
    JFEM_CUDA(cudaMemsetAsync(b.p, 0, b.bytes(), h->stream));
    JFEM_CUDA(cudaMemsetAsync(x.p, 0, x.bytes(), h->stream));
    JFEM_CUDA(cudaMemsetAsync(p.p, 0, p.bytes(), h->stream));
    MatBase mat;
    fill_material(h, mat);
    for (size_t c = 0; c + 1 < h->colour_ptr.size(); c++) {
        const long long e0 = h->colour_ptr[c], ne = h->colour_ptr[c + 1] - e0;
        if (!ne) continue;
        const unsigned g = (unsigned)((ne + 63) / 64);
        mass_assemble_kernel<<<g, 64, 0, h->stream>>>(nnpe, ne, h->colour_elems.p, e0, h->dconn.p, h->coords.p, (const long long *)h->nadj_ptr.p, h->eblk.p, M.p);
        recover_rhs_kernel<<<g, 64, 0, h->stream>>>(nnpe, ne, h->colour_elems.p, e0, h->dconn.p, h->coords.p, du, mat, (const long long *)h->e2i.p, field, b.p);
        h->total_launches += 2;
    }
    JFEM_CUDA(cudaGetLastError());
    const unsigned gn = (unsigned)((nn + 127) / 128), g6 = (unsigned)((6 * nn + 255) / 256);
    mass_diag_kernel<<<gn, 128, 0, h->stream>>>(nn, (const long long *)h->nadj_ptr.p, h->nadj.p, M.p, dinv.p);
    // Jacobi-PCG, 6 right-hand sides with their own scalars; x0 = 0, r0 = b
    JFEM_CUDA(cudaMemcpyAsync(r.p, b.p, b.bytes(), cudaMemcpyDeviceToDevice, h->stream));
    double hs[24], rz[6], rz0[6], zero6[6] = {0, 0, 0, 0, 0, 0};
    double *d_rz = sc.p, *d_pAp = sc.p + 6, *d_alpha = sc.p + 12, *d_beta = sc.p + 18;
    JFEM_CUDA(cudaMemcpyAsync(d_beta, zero6, sizeof zero6, cudaMemcpyHostToDevice, h->stream));
    dot6_kernel<<<1, 256, 0, h->stream>>>(nn, r.p, r.p, dinv.p, d_rz);
    JFEM_CUDA(cudaMemcpyAsync(rz, d_rz, sizeof rz, cudaMemcpyDeviceToHost, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    for (int q = 0; q < 6; q++) rz0[q] = rz[q];
    int it = 0;
    for (; it < 500; it++) {
        bool done = true;
        for (int q = 0; q < 6; q++) if (rz[q] > 1e-28 * rz0[q] && rz[q] > 0) done = false;     // ||r||_{D^-1} <= 1e-14 ||b||_{D^-1}
        if (done) break;
        pcg6_dir_kernel<<<g6, 256, 0, h->stream>>>(nn, d_beta, dinv.p, r.p, p.p);
        mass_spmv6_kernel<<<gn, 128, 0, h->stream>>>(nn, (const long long *)h->nadj_ptr.p, h->nadj.p, M.p, p.p, Ap.p);
        dot6_kernel<<<1, 256, 0, h->stream>>>(nn, p.p, Ap.p, nullptr, d_pAp);
        JFEM_CUDA(cudaMemcpyAsync(hs, d_pAp, 6 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        JFEM_CUDA(cudaStreamSynchronize(h->stream));
        double alpha[6];
        for (int q = 0; q < 6; q++) alpha[q] = hs[q] != 0 ? rz[q] / hs[q] : 0.0;
        JFEM_CUDA(cudaMemcpyAsync(d_alpha, alpha, sizeof alpha, cudaMemcpyHostToDevice, h->stream));
        pcg6_update_kernel<<<g6, 256, 0, h->stream>>>(nn, d_alpha, p.p, Ap.p, x.p, r.p);
        dot6_kernel<<<1, 256, 0, h->stream>>>(nn, r.p, r.p, dinv.p, d_rz);
        double rzn[6], beta[6];
        JFEM_CUDA(cudaMemcpyAsync(rzn, d_rz, sizeof rzn, cudaMemcpyDeviceToHost, h->stream));
        JFEM_CUDA(cudaStreamSynchronize(h->stream));
        for (int q = 0; q < 6; q++) { beta[q] = rz[q] != 0 ? rzn[q] / rz[q] : 0.0; rz[q] = rzn[q]; }
        JFEM_CUDA(cudaMemcpyAsync(d_beta, beta, sizeof beta, cudaMemcpyHostToDevice, h->stream));
        h->total_launches += 5;
    }
    JFEM_CUDA(cudaGetLastError());
    JFEM_CUDA(cudaMemcpyAsync(out, x.p, 6 * (size_t)nn * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    M.release(); b.release(); x.release(); r.release(); p.release(); Ap.release(); dinv.release(); sc.release();
    if (it >= 500) { jfem_set_error("nodal recovery: mass-matrix PCG did not converge in 500 iterations"); return JFEM_ESTATE; }
    return JFEM_OK;
}

int jfem_csr_penalty_bc(jfem_handle *h, double scale, double *rhs, double *penalty_out, int on_device) {
    CHECK_H(h);
    if (!h->vals_valid) { jfem_set_error("jfem_assemble_csr has not been called"); return JFEM_ESTATE; }
    const long long nr = h->n_dofs();
    DevBuf<double> mx, drhs;
    JFEM_TRY(mx.alloc(1));
    absmax_kernel<<<1, 256, 0, h->stream>>>((long long)h->vals.n, h->vals.p, mx.p);
    double *dr = rhs;
    if (rhs && !on_device) {
        JFEM_TRY(drhs.alloc((size_t)nr));
        JFEM_CUDA(cudaMemcpyAsync(drhs.p, rhs, nr * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        dr = drhs.p;
    }
    penalty_kernel<<<(unsigned)((nr + 127) / 128), 128, 0, h->stream>>>(nr, (const long long *)h->rowptr.p, h->colind.p, h->fixed.p, h->prescribed.p, mx.p, scale,
                                                                         h->vals.p, dr);
    JFEM_CUDA(cudaGetLastError());
    h->total_launches += 2;
    if (rhs && !on_device) JFEM_CUDA(cudaMemcpyAsync(rhs, dr, nr * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    double m = 0;
    JFEM_CUDA(cudaMemcpyAsync(&m, mx.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    JFEM_CUDA(cudaStreamSynchronize(h->stream));
    if (penalty_out) *penalty_out = scale * m;
    mx.release(); drhs.release();
    return JFEM_OK;
}

}  // extern "C"

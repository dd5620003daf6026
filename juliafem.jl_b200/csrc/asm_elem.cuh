// asm_elem.cuh -- element tangent by columns with the geometry shared between the columns (assembled path).
//
// assemble_element! of the reference (src/problems_elasticity.jl:245-409) builds, per Gauss point, BL from grad N and
// accumulates Km += w BL' D BL (+ Kg).  Here the same sum is organised for one warp per element:
//   gp_geometry   one lane per Gauss point: J = sum_k dN_k (x) X_k, inv(J), grad N_k = inv(J) dN_k, w = weight det(J)
//                 (src/basis/math.jl:47-54,198-201), written once to shared memory;
//   gp_column     one lane per column (node l, component cj) of Ke: the tangent functor of elem.cuh is applied to the
//                 gradient of the unit displacement e_cj N_l, G = e_cj (x) grad N_l, and the resulting stress increment is
//                 contracted with grad N_k of every row node: Ke[(k,i),(l,cj)] += w sum_j dP[i][j] grad N_k[j].
// Every column reuses the geometry of its element (the column-per-thread kernel recomputes it 3*nnpe times).  The functions are
// __host__ __device__: tests/hostcheck replays them on the CPU against the oracle's element matrices.
#pragma once
#include "elem.cuh"

namespace jf {

template <int NNPE> struct ElemRule;
template <> struct ElemRule<10> { static constexpr int NGP = 4; };   // GLTET4  (src/quadrature/gltet.jl:18-25)
template <> struct ElemRule<8> { static constexpr int NGP = 8; };    // GLHEX8  (src/quadrature/glquad.jl:6-44, quaddata.jl:4-5)
template <> struct ElemRule<4> { static constexpr int NGP = 1; };    // GLTET1  (src/quadrature/gltet.jl:7-11)

// dN[k][a] = dN_k / dxi_a at Gauss point g (same point order as the *_general kernels of elem.cuh, so that the state
// index elem * NGP + g means the same point); returns the quadrature weight.
template <int NNPE>
JF_HD double ref_derivs(int g, double (&dN)[NNPE][3]) {
    if constexpr (NNPE == 10) {
        // barycentric form: dN_a/dL_b = delta_ab (4 L_a - 1), dN_ab/dL_a = 4 L_b; xi_c = L_(c+1), L_0 = 1 - sum xi
        const int s = (g + 1) & 3;   // GLTET4 point g has L_s = A, the other three = B
        double L[4], dL[10][4];
        JF_UNROLL for (int b = 0; b < 4; b++) L[b] = (b == s) ? T10_A : T10_B;
        JF_UNROLL for (int k = 0; k < 10; k++) JF_UNROLL for (int b = 0; b < 4; b++) dL[k][b] = 0.0;
        JF_UNROLL for (int a = 0; a < 4; a++) dL[a][a] = 4.0 * L[a] - 1.0;
        JF_UNROLL for (int a = 0; a < 4; a++) JF_UNROLL for (int b = a + 1; b < 4; b++) {
            dL[t10_edge(a, b)][a] = 4.0 * L[b];
            dL[t10_edge(a, b)][b] = 4.0 * L[a];
        }
        JF_UNROLL for (int k = 0; k < 10; k++) JF_UNROLL for (int c = 0; c < 3; c++) dN[k][c] = dL[k][c + 1] - dL[k][0];
        return 1.0 / 24.0;
    } else if constexpr (NNPE == 8) {
        const double u = (g & 1) ? HEX_GA : -HEX_GA, v = (g & 2) ? HEX_GA : -HEX_GA, w = (g & 4) ? HEX_GA : -HEX_GA;
        JF_UNROLL for (int k = 0; k < 8; k++) {
            const double a = (k == 1 || k == 2 || k == 5 || k == 6) ? 1.0 : -1.0;
            const double b = (k == 2 || k == 3 || k == 6 || k == 7) ? 1.0 : -1.0;
            const double c = (k >= 4) ? 1.0 : -1.0;
            dN[k][0] = 0.125 * a * (1.0 + b * v) * (1.0 + c * w);
            dN[k][1] = 0.125 * b * (1.0 + a * u) * (1.0 + c * w);
            dN[k][2] = 0.125 * c * (1.0 + a * u) * (1.0 + b * v);
        }
        return 1.0;
    } else {
        JF_UNROLL for (int c = 0; c < 3; c++) dN[0][c] = -1.0;
        JF_UNROLL for (int a = 1; a < 4; a++) JF_UNROLL for (int c = 0; c < 3; c++) dN[a][c] = (a - 1 == c) ? 1.0 : 0.0;
        return 1.0 / 6.0;
    }
}

// Geometry of Gauss point g: gN[3 k + j] = d N_k / d x_j, returns w = weight * det(J).  X(k, c) = coordinate c of element node k.
template <int NNPE, class XFLD>
JF_HD double gp_geometry(int g, const XFLD &X, double *gN) {
    double dN[NNPE][3];
    const double wq = ref_derivs<NNPE>(g, dN);
    double J[3][3], iJ[3][3];   // J[a][c] = dx_c / dxi_a
    JF_UNROLL for (int a = 0; a < 3; a++) JF_UNROLL for (int c = 0; c < 3; c++) J[a][c] = 0.0;
    JF_UNROLL for (int k = 0; k < NNPE; k++) {
        const double x0 = X(k, 0), x1 = X(k, 1), x2 = X(k, 2);
        JF_UNROLL for (int a = 0; a < 3; a++) { J[a][0] += dN[k][a] * x0; J[a][1] += dN[k][a] * x1; J[a][2] += dN[k][a] * x2; }
    }
    const double det = inv3x3(J, iJ);   // d/dx_j = sum_a iJ[j][a] d/dxi_a
    JF_UNROLL for (int k = 0; k < NNPE; k++) JF_UNROLL for (int j = 0; j < 3; j++)
        gN[3 * k + j] = iJ[j][0] * dN[k][0] + iJ[j][1] * dN[k][1] + iJ[j][2] * dN[k][2];
    return det * wq;
}

// Gu[3 i + j] = d u_i / d x_j at the Gauss point whose gradients are gN (linearisation point of the tangent functors)
template <int NNPE, class FLD>
JF_HD void gp_grad(const FLD &U, const double *gN, double *Gu) {
    JF_UNROLL for (int q = 0; q < 9; q++) Gu[q] = 0.0;
    JF_UNROLL for (int k = 0; k < NNPE; k++) JF_UNROLL for (int i = 0; i < 3; i++) {
        const double u = U(k, i);
        JF_UNROLL for (int j = 0; j < 3; j++) Gu[3 * i + j] += u * gN[3 * k + j];
    }
}

// Contribution of one Gauss point to column (l, cj) of the element tangent: acc[k][i] += w sum_j dP[i][j] gN[k][j].
// pt must already hold the element's parameters (Pt::load).  Gu is read by the two-field (tangent) functors only.
template <int NNPE, class Pt>
JF_HD bool gp_column(const Pt &pt, long long gp, const double *gN, double w, const double *Gu, int l, int cj, double (&acc)[NNPE][3]) {
    constexpr int NF = Pt::NF;
    double G[NF][3][3], P[3][3];
    JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) {
        G[0][i][j] = (i == cj) ? gN[3 * l + j] : 0.0;
        if (NF == 2) G[NF - 1][i][j] = Gu[3 * i + j];
    }
    const bool ok = pt.eval(gp, G, P);
    JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) P[i][j] *= w;
    JF_UNROLL for (int k = 0; k < NNPE; k++) {
        const double g0 = gN[3 * k], g1 = gN[3 * k + 1], g2 = gN[3 * k + 2];
        JF_UNROLL for (int i = 0; i < 3; i++) acc[k][i] += P[i][0] * g0 + P[i][1] * g1 + P[i][2] * g2;
    }
    return ok;
}

}  // namespace jf

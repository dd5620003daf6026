// elem.cuh -- per-element device arithmetic (one thread = one element), fp64, registers only.
//
// What is computed is the reference's element integration (src/problems_elasticity.jl:245-409):
//   J = sum_i dN_i (x) X_i, grad_i = inv(J).dN_i, grad(u) = sum_k u_k (x) grad_k   (src/basis/math.jl:47-54,198-201,249-255)
//   f_k += w * P . grad_k  with w = weight*det(J)                                   (problems_elasticity.jl:248,407-409)
// but never through the 6 x ndof BL matrix: gradients are taken in reference space and pushed through
// inv(J) (3x3), and the Tet10 shape-function derivatives are used in barycentric form
//   dN_a/dL_b = delta_ab (4 L_a - 1),  dN_ab/dL_a = 4 L_b      (same polynomials as lagrange_generated.jl:275-282)
// which turns the 10x3 derivative table at the four GLTET4 points (src/quadrature/gltet.jl:18-25) into
// "base + one extra term" updates.  Point functors (`Pt`) supply the constitutive map grad(u) -> P.
#pragma once
#include <cuda_runtime.h>

namespace jf {

#define JF_UNROLL _Pragma("unroll")
#define JF_HD __host__ __device__ __forceinline__   // host instantiation is used by the CPU-side layout check only (tests/hostcheck)

// GLTET4 abscissae (5+3*sqrt(5))/20 and (5-sqrt(5))/20
#define T10_A 0.58541019662496845
#define T10_B 0.13819660112501052

__host__ __device__ constexpr int t10_edge(int a, int b) {
    // mid-edge node of vertices a != b; order 4=(0-1) 5=(1-2) 6=(0-2) 7=(0-3) 8=(1-3) 9=(2-3)  (lagrange_generated.jl:261-262)
    return (a < b ? a : b) == 0 ? ((a < b ? b : a) == 1 ? 4 : ((a < b ? b : a) == 2 ? 6 : 7))
         : (a < b ? a : b) == 1 ? ((a < b ? b : a) == 2 ? 5 : 8)
                                : 9;
}

JF_HD double inv3x3(const double (&J)[3][3], double (&iJ)[3][3]) {
    double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    double r = 1.0 / det;
    iJ[0][0] = c00 * r; iJ[1][0] = c01 * r; iJ[2][0] = c02 * r;
    iJ[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * r;
    iJ[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * r;
    iJ[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * r;
    iJ[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * r;
    iJ[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * r;
    iJ[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * r;
    return det;
}

// field accessor over shared memory: value of component c at local element node k
struct SField {
    const double *base;   // 3 doubles per patch node
    const int *n;         // local node index of the element's nodes
    JF_HD double operator()(int k, int c) const { return base[3 * n[k] + c]; }
};

// ------------------------------------------------------------------------------------------------
// Point functors: P[i][j] from displacement gradients.  G[f][i][j] = d(field f)_i / dx_j.
// ------------------------------------------------------------------------------------------------

// Material parameters: homogeneous (pe == nullptr) or per element (pe = SoA [param][internal element], like the
// E_vec / nu_vec arrays of ext/JuliaFEMCUDAExt.jl:135-141).  load() is called once per element.
struct MatBase {
    double la, mu, sy, H;
    const double *pe;
    long long pe_n;
    JF_HD void load(long long elem) {
        if (pe) {
            const double E = pe[elem], nu = pe[pe_n + elem];
            la = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu));   // src/materials/linear_elastic.jl:82
            mu = E / (2.0 * (1.0 + nu));                     // :97
            sy = pe[2 * pe_n + elem]; H = pe[3 * pe_n + elem];
        }
    }
};

// sigma = la tr(eps) I + 2 mu eps  (src/materials/linear_elastic.jl:136-158; problems_elasticity.jl:313-332)
struct PtLinear : MatBase {
    static constexpr int NF = 1;
    JF_HD bool eval(long long, const double (&G)[1][3][3], double (&P)[3][3]) const {
        double tr = la * (G[0][0][0] + G[0][1][1] + G[0][2][2]);
        double s01 = mu * (G[0][0][1] + G[0][1][0]), s12 = mu * (G[0][1][2] + G[0][2][1]), s02 = mu * (G[0][0][2] + G[0][2][0]);
        P[0][0] = tr + 2 * mu * G[0][0][0]; P[1][1] = tr + 2 * mu * G[0][1][1]; P[2][2] = tr + 2 * mu * G[0][2][2];
        P[0][1] = P[1][0] = s01; P[1][2] = P[2][1] = s12; P[0][2] = P[2][0] = s02;
        return true;
    }
};

JF_HD void sym6_to_33(const double (&s)[6], double (&S)[3][3]) {
    S[0][0] = s[0]; S[1][1] = s[1]; S[2][2] = s[2];
    S[0][1] = S[1][0] = s[3]; S[1][2] = S[2][1] = s[4]; S[0][2] = S[2][0] = s[5];
}

// Compressible Neo-Hookean, psi = mu/2 (I1-3) - mu ln J + la/2 ln^2 J  (src/materials/neo_hookean.jl:129-143).
// S = 2 dpsi/dC = mu (I - C^-1) + la lnJ C^-1 ;  DD = 4 d2psi/dC2 = la Ci(x)Ci + 2(mu - la lnJ) I_{Ci}
// (closed form of the Tensors.hessian call at neo_hookean.jl:222).  Total Lagrangian: P = F S,
// dP = dF S + F (DD : sym(F' dF))  -- material + geometric stiffness (problems_elasticity.jl:270-289,378-404).
struct NHCommon : MatBase {
    JF_HD bool kin(const double (&Gu)[3][3], double (&F)[3][3], double (&Ci)[3][3], double &lnJ) const {
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) F[i][j] = Gu[i][j] + (i == j ? 1.0 : 0.0);
        double C[3][3];
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) C[i][j] = F[0][i] * F[0][j] + F[1][i] * F[1][j] + F[2][i] * F[2][j];
        double d = inv3x3(C, Ci);
        lnJ = 0.5 * log(d);
        return d > 0.0;
    }
};

struct PtNHResidual : NHCommon {
    static constexpr int NF = 1;
    JF_HD bool eval(long long, const double (&G)[1][3][3], double (&P)[3][3]) const {
        double F[3][3], Ci[3][3], lnJ;
        bool ok = kin(G[0], F, Ci, lnJ);
        double S[3][3];
        double c = la * lnJ - mu;
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) S[i][j] = (i == j ? mu : 0.0) + c * Ci[i][j];
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) P[i][j] = F[i][0] * S[0][j] + F[i][1] * S[1][j] + F[i][2] * S[2][j];
        return ok;
    }
};

struct PtNHTangent : NHCommon {   // field 0 = v (direction), field 1 = u (linearisation point)
    static constexpr int NF = 2;
    JF_HD bool eval(long long, const double (&G)[2][3][3], double (&P)[3][3]) const {
        double F[3][3], Ci[3][3], lnJ;
        bool ok = kin(G[1], F, Ci, lnJ);
        double c = la * lnJ - mu;
        double S[3][3], dE[3][3], A[3][3];
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) S[i][j] = (i == j ? mu : 0.0) + c * Ci[i][j];
        // A = F' dF ; dE = sym(A)
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) A[i][j] = F[0][i] * G[0][0][j] + F[1][i] * G[0][1][j] + F[2][i] * G[0][2][j];
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) dE[i][j] = 0.5 * (A[i][j] + A[j][i]);
        // dS = la (Ci:dE) Ci + 2 (mu - la lnJ) Ci dE Ci
        double tr = 0;
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) tr += Ci[i][j] * dE[i][j];
        double B[3][3], dS[3][3];
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) B[i][j] = Ci[i][0] * dE[0][j] + Ci[i][1] * dE[1][j] + Ci[i][2] * dE[2][j];
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++)
            dS[i][j] = la * tr * Ci[i][j] - 2.0 * c * (B[i][0] * Ci[0][j] + B[i][1] * Ci[1][j] + B[i][2] * Ci[2][j]);
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++)
            P[i][j] = G[0][i][0] * S[0][j] + G[0][i][1] * S[1][j] + G[0][i][2] * S[2][j] + F[i][0] * dS[0][j] + F[i][1] * dS[1][j] + F[i][2] * dS[2][j];
        return ok;
    }
};

// St. Venant-Kirchhoff = the reference's classic finite-strain path (props.finite_strain = true, src/problems_elasticity.jl:
// 255-332): Hooke's D applied to the Green-Lagrange strain E = 1/2 (grad u + grad u' + grad u' grad u), S = la tr(E) I + 2 mu E,
// Total Lagrangian P = F S.  Tangent: dP = F (D : sym(F' dF))  [BL' D BL, :270-289,370-375]  (+ dF S, the geometric
// stiffness Kg of :378-404, only when props.geometric_stiffness is set -- `geo`).
struct StVKCommon : MatBase {
    int geo;
    JF_HD void kin(const double (&Gu)[3][3], double (&F)[3][3], double (&S)[3][3]) const {
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) F[i][j] = Gu[i][j] + (i == j ? 1.0 : 0.0);
        double E[3][3];
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++)
            E[i][j] = 0.5 * (Gu[i][j] + Gu[j][i] + Gu[0][i] * Gu[0][j] + Gu[1][i] * Gu[1][j] + Gu[2][i] * Gu[2][j]);
        const double tr = la * (E[0][0] + E[1][1] + E[2][2]);
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) S[i][j] = 2.0 * mu * E[i][j] + (i == j ? tr : 0.0);
    }
};

struct PtStVKResidual : StVKCommon {
    static constexpr int NF = 1;
    JF_HD bool eval(long long, const double (&G)[1][3][3], double (&P)[3][3]) const {
        double F[3][3], S[3][3];
        kin(G[0], F, S);
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) P[i][j] = F[i][0] * S[0][j] + F[i][1] * S[1][j] + F[i][2] * S[2][j];
        return true;
    }
};

struct PtStVKTangent : StVKCommon {   // field 0 = v (direction), field 1 = u (linearisation point)
    static constexpr int NF = 2;
    JF_HD bool eval(long long, const double (&G)[2][3][3], double (&P)[3][3]) const {
        double F[3][3], S[3][3], A[3][3], dS[3][3];
        kin(G[1], F, S);
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) A[i][j] = F[0][i] * G[0][0][j] + F[1][i] * G[0][1][j] + F[2][i] * G[0][2][j];
        const double tr = la * (A[0][0] + A[1][1] + A[2][2]);
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) dS[i][j] = mu * (A[i][j] + A[j][i]) + (i == j ? tr : 0.0);
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) {
            double v = F[i][0] * dS[0][j] + F[i][1] * dS[1][j] + F[i][2] * dS[2][j];
            if (geo) v += G[0][i][0] * S[0][j] + G[0][i][1] * S[1][j] + G[0][i][2] * S[2][j];
            P[i][j] = v;
        }
        return true;
    }
};

// J2 plasticity with linear kinematic hardening, radial return (src/materials/perfect_plasticity.jl:247-341).
// State per Gauss point, SoA: st[s * n_gp + gp], s = 0..12 (eps_p 11,22,33,12,23,13 ; alpha ; kappa).
struct PPCommon : MatBase {
    const double *st_old;
    long long n_gp;
    // returns true if plastic; s = stress (tensor comps), n = flow direction s_trial/q, dl = plastic multiplier
    JF_HD bool ret_map(long long gp, const double (&e)[6], double (&s)[6], double (&n)[6], double &dl) const {
        double ee[6], al[6];
        JF_UNROLL for (int i = 0; i < 6; i++) { ee[i] = e[i] - st_old[i * n_gp + gp]; al[i] = st_old[(6 + i) * n_gp + gp]; }
        double tr = la * (ee[0] + ee[1] + ee[2]);
        JF_UNROLL for (int i = 0; i < 3; i++) s[i] = tr + 2 * mu * ee[i];
        JF_UNROLL for (int i = 3; i < 6; i++) s[i] = 2 * mu * ee[i];
        double sd[6];
        JF_UNROLL for (int i = 0; i < 6; i++) sd[i] = s[i] - al[i];
        double m = (sd[0] + sd[1] + sd[2]) / 3.0;
        sd[0] -= m; sd[1] -= m; sd[2] -= m;
        double nn = sd[0] * sd[0] + sd[1] * sd[1] + sd[2] * sd[2] + 2 * (sd[3] * sd[3] + sd[4] * sd[4] + sd[5] * sd[5]);
        double q = sqrt(3.0 / 2.0) * sqrt(nn);
        double f = q - sy;
        if (f <= 0.0) { dl = 0.0; JF_UNROLL for (int i = 0; i < 6; i++) n[i] = 0.0; return false; }
        JF_UNROLL for (int i = 0; i < 6; i++) n[i] = sd[i] / q;
        dl = f / (2 * mu + (2.0 / 3.0) * H);
        JF_UNROLL for (int i = 0; i < 6; i++) s[i] -= 2 * mu * dl * n[i];
        return true;
    }
};

JF_HD void strain6(const double (&G)[3][3], double (&e)[6]) {
    e[0] = G[0][0]; e[1] = G[1][1]; e[2] = G[2][2];
    e[3] = 0.5 * (G[0][1] + G[1][0]); e[4] = 0.5 * (G[1][2] + G[2][1]); e[5] = 0.5 * (G[0][2] + G[2][0]);
}

struct PtPPResidual : PPCommon {   // also writes the trial state (commit-on-convergence, abstract_material.jl:203-207)
    static constexpr int NF = 1;
    double *st_new;
    JF_HD bool eval(long long gp, const double (&G)[1][3][3], double (&P)[3][3]) const {
        double e[6], s[6], n[6], dl;
        strain6(G[0], e);
        bool pl = ret_map(gp, e, s, n, dl);
        if (st_new) {
            JF_UNROLL for (int i = 0; i < 6; i++) {
                st_new[i * n_gp + gp] = st_old[i * n_gp + gp] + dl * n[i];
                st_new[(6 + i) * n_gp + gp] = st_old[(6 + i) * n_gp + gp] + (2.0 / 3.0) * H * dl * n[i];
            }
            st_new[12 * n_gp + gp] = st_old[12 * n_gp + gp] + dl;
        }
        (void)pl;
        sym6_to_33(s, P);
        return true;
    }
};

struct PtPPTangent : PPCommon {   // field 0 = v, field 1 = u;  dP = DD_ep : sym(grad v)
    static constexpr int NF = 2;
    JF_HD bool eval(long long gp, const double (&G)[2][3][3], double (&P)[3][3]) const {
        double e[6], s[6], n[6], dl, de[6], ds[6];
        strain6(G[1], e);
        bool pl = ret_map(gp, e, s, n, dl);
        strain6(G[0], de);
        double tr = la * (de[0] + de[1] + de[2]);
        JF_UNROLL for (int i = 0; i < 3; i++) ds[i] = tr + 2 * mu * de[i];
        JF_UNROLL for (int i = 3; i < 6; i++) ds[i] = 2 * mu * de[i];
        if (pl) {   // DD = DD_e - 4 mu^2/(2 mu + 2H/3) n (x) n   (perfect_plasticity.jl:337)
            double nd = n[0] * de[0] + n[1] * de[1] + n[2] * de[2] + 2 * (n[3] * de[3] + n[4] * de[4] + n[5] * de[5]);
            double c = 4 * mu * mu / (2 * mu + (2.0 / 3.0) * H) * nd;
            JF_UNROLL for (int i = 0; i < 6; i++) ds[i] -= c * n[i];
        }
        sym6_to_33(ds, P);
        return true;
    }
};

// ------------------------------------------------------------------------------------------------
// Tet10, isoparametric (valid for curved elements), GLTET4.  F[f] are the NF fields to differentiate,
// X the coordinates; out(k, v0, v1, v2) receives the 3 components of node k of the element vector.  Returns false on invalid deformation.
// ------------------------------------------------------------------------------------------------
template <class Pt, class FLD, class OUT>
JF_HD bool tet10_general(const Pt &pt0, long long elem, const FLD (&F)[Pt::NF], const FLD &X, OUT &&out) {
    constexpr int NF = Pt::NF;
    Pt pt = pt0;
    pt.load(elem);
    const double c1 = 4.0 * T10_B - 1.0, c4b = 4.0 * T10_B, k4 = 4.0 * (T10_A - T10_B);
    double bx[4][3], bu[NF][4][3];
    JF_UNROLL for (int a = 0; a < 4; a++) JF_UNROLL for (int c = 0; c < 3; c++) {
        double sx = 0;
        JF_UNROLL for (int b = 0; b < 4; b++) if (b != a) sx += X(t10_edge(a, b), c);
        bx[a][c] = c1 * X(a, c) + c4b * sx;
        JF_UNROLL for (int f = 0; f < NF; f++) {
            double su = 0;
            JF_UNROLL for (int b = 0; b < 4; b++) if (b != a) su += F[f](t10_edge(a, b), c);
            bu[f][a][c] = c1 * F[f](a, c) + c4b * su;
        }
    }
    double ST[4][3], fo[10][3];
    JF_UNROLL for (int a = 0; a < 4; a++) JF_UNROLL for (int c = 0; c < 3; c++) ST[a][c] = 0.0;
    bool ok = true;
    JF_UNROLL for (int g = 0; g < 4; g++) {
        const int s = (g + 1) & 3;   // GLTET4 point g has L_s = A, the other three = B
        double q0[3], J[3][3], iJ[3][3];
        JF_UNROLL for (int c = 0; c < 3; c++) q0[c] = bx[0][c] + k4 * X(s == 0 ? 0 : t10_edge(0, s), c);
        JF_UNROLL for (int a = 1; a < 4; a++) JF_UNROLL for (int c = 0; c < 3; c++)
            J[a - 1][c] = bx[a][c] + k4 * X(a == s ? a : t10_edge(a, s), c) - q0[c];     // J[a][b] = dx_b/dxi_a
        double det = inv3x3(J, iJ);
        double w = det * (1.0 / 24.0);
        double G[NF][3][3];
        JF_UNROLL for (int f = 0; f < NF; f++) {
            double h0[3], Hh[3][3];
            JF_UNROLL for (int c = 0; c < 3; c++) h0[c] = bu[f][0][c] + k4 * F[f](s == 0 ? 0 : t10_edge(0, s), c);
            JF_UNROLL for (int a = 1; a < 4; a++) JF_UNROLL for (int c = 0; c < 3; c++)
                Hh[a - 1][c] = bu[f][a][c] + k4 * F[f](a == s ? a : t10_edge(a, s), c) - h0[c];  // H[a][i] = du_i/dxi_a
            JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++)
                G[f][i][j] = iJ[j][0] * Hh[0][i] + iJ[j][1] * Hh[1][i] + iJ[j][2] * Hh[2][i];
        }
        double P[3][3];
        ok &= pt.eval(elem * 4 + g, G, P);
        // T_b[i] = w * sum_j P_ij gradL_b[j];  gradL_{c+1}[j] = iJ[j][c], gradL_0 = -sum
        double T[4][3];
        JF_UNROLL for (int i = 0; i < 3; i++) {
            JF_UNROLL for (int c = 0; c < 3; c++) T[c + 1][i] = w * (P[i][0] * iJ[0][c] + P[i][1] * iJ[1][c] + P[i][2] * iJ[2][c]);
            T[0][i] = -(T[1][i] + T[2][i] + T[3][i]);
        }
        JF_UNROLL for (int a = 0; a < 4; a++) JF_UNROLL for (int i = 0; i < 3; i++) ST[a][i] += T[a][i];
        JF_UNROLL for (int i = 0; i < 3; i++) fo[s][i] = k4 * T[s][i];
        JF_UNROLL for (int a = 0; a < 4; a++) if (a != s) {
            const int e = t10_edge(a, s);
            const bool first = ((s + 3) & 3) < ((a + 3) & 3);   // GP order visits s = 1,2,3,0
            JF_UNROLL for (int i = 0; i < 3; i++) fo[e][i] = first ? k4 * T[a][i] : fo[e][i] + k4 * T[a][i];
        }
    }
    JF_UNROLL for (int a = 0; a < 4; a++) out(a, fo[a][0] + c1 * ST[a][0], fo[a][1] + c1 * ST[a][1], fo[a][2] + c1 * ST[a][2]);
    JF_UNROLL for (int a = 0; a < 4; a++) JF_UNROLL for (int b = a + 1; b < 4; b++) {
        const int e = t10_edge(a, b);
        out(e, fo[e][0] + c4b * (ST[a][0] + ST[b][0]), fo[e][1] + c4b * (ST[a][1] + ST[b][1]), fo[e][2] + c4b * (ST[a][2] + ST[b][2]));
    }
    return ok;
}

// ------------------------------------------------------------------------------------------------
// Tet10 with affine geometry + linear elasticity: grad(u) is linear over the element, so the GLTET4 sum
// (exact for the quadratic integrand) is replaced by its closed form on vertex values:
//   G_c   = grad u at vertex c = 4 (u_c (x) g_c + sum_{a!=c} u_ac (x) g_a) - sum_a u_a (x) g_a ,  g_a = grad L_a
//   f_a   = V/20 (4 sigma_a - S) g_a ,  f_ab = V/5 ((S + sigma_b) g_a + (S + sigma_a) g_b) ,  S = sum_c sigma_c
// Only the 4 vertex coordinates are read.  The constants are folded so that no stand-alone scaling multiplies remain:
// with h = 4 g and t_c = V/20 sigma_c (Lame constants pre-scaled by V/20), T = sum t_c:
//   G_c = u_c (x) h_c + sum_{a!=c} u_ac (x) h_a - W ,  f_a = (t_a - T/4) h_a ,  f_ab = (T + t_b) h_a + (T + t_a) h_b
// (~520 fp64 instructions per element).
// ------------------------------------------------------------------------------------------------
template <class FLD, class XFLD, class OUT>
JF_HD void tet10_affine_linear(double la, double mu, const FLD &U, const XFLD &X, OUT &&out) {
    double J[3][3];
    JF_UNROLL for (int a = 0; a < 3; a++) JF_UNROLL for (int c = 0; c < 3; c++) J[a][c] = X(a + 1, c) - X(0, c);
    // cofactors: cf[j][a] * (1/det) = d L_{a+1} / d x_j
    double cf[3][3];
    cf[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1]; cf[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2]; cf[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    cf[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2]; cf[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0]; cf[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    cf[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1]; cf[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2]; cf[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double det = J[0][0] * cf[0][0] + J[0][1] * cf[1][0] + J[0][2] * cf[2][0];
    const double r = 1.0 / det, r4 = 4.0 * r;
    double W[3][3];   // sum_a u_a (x) g_a
    {
        double g[4][3];
        JF_UNROLL for (int j = 0; j < 3; j++) {
            g[1][j] = cf[j][0] * r; g[2][j] = cf[j][1] * r; g[3][j] = cf[j][2] * r;
            g[0][j] = -(g[1][j] + g[2][j] + g[3][j]);
        }
        JF_UNROLL for (int i = 0; i < 3; i++) {
            const double u0 = U(0, i), u1 = U(1, i), u2 = U(2, i), u3 = U(3, i);
            JF_UNROLL for (int j = 0; j < 3; j++) W[i][j] = u0 * g[0][j] + u1 * g[1][j] + u2 * g[2][j] + u3 * g[3][j];
        }
    }
    double h[4][3];   // 4 grad L_a
    JF_UNROLL for (int j = 0; j < 3; j++) {
        h[1][j] = cf[j][0] * r4; h[2][j] = cf[j][1] * r4; h[3][j] = cf[j][2] * r4;
        h[0][j] = -(h[1][j] + h[2][j] + h[3][j]);
    }
    const double v20 = det * (1.0 / 120.0), las = la * v20, mus = mu * v20, mus2 = mus + mus;
    double t[4][6];   // V/20 * vertex stresses, 11 22 33 12 23 13
    JF_UNROLL for (int c = 0; c < 4; c++) {
        double G[3][3];
        JF_UNROLL for (int i = 0; i < 3; i++) {
            const double uc = U(c, i);
            JF_UNROLL for (int j = 0; j < 3; j++) G[i][j] = uc * h[c][j] - W[i][j];
        }
        JF_UNROLL for (int a = 0; a < 4; a++) if (a != c) JF_UNROLL for (int i = 0; i < 3; i++) {
            const double ue = U(t10_edge(a, c), i);
            JF_UNROLL for (int j = 0; j < 3; j++) G[i][j] += ue * h[a][j];
        }
        const double tr = las * (G[0][0] + G[1][1] + G[2][2]);
        t[c][0] = tr + mus2 * G[0][0]; t[c][1] = tr + mus2 * G[1][1]; t[c][2] = tr + mus2 * G[2][2];
        t[c][3] = mus * (G[0][1] + G[1][0]); t[c][4] = mus * (G[1][2] + G[2][1]); t[c][5] = mus * (G[0][2] + G[2][0]);
    }
    double T[6], Tq[6];
    JF_UNROLL for (int q = 0; q < 6; q++) { T[q] = (t[0][q] + t[1][q]) + (t[2][q] + t[3][q]); Tq[q] = 0.25 * T[q]; }
    auto mulsym = [](const double (&s)[6], const double (&v)[3], double (&o)[3]) {
        o[0] = s[0] * v[0] + s[3] * v[1] + s[5] * v[2];
        o[1] = s[3] * v[0] + s[1] * v[1] + s[4] * v[2];
        o[2] = s[5] * v[0] + s[4] * v[1] + s[2] * v[2];
    };
    JF_UNROLL for (int a = 0; a < 4; a++) {
        double mm[6], o[3];
        JF_UNROLL for (int q = 0; q < 6; q++) mm[q] = t[a][q] - Tq[q];
        mulsym(mm, h[a], o);
        out(a, o[0], o[1], o[2]);
    }
    JF_UNROLL for (int a = 0; a < 4; a++) JF_UNROLL for (int q = 0; q < 6; q++) t[a][q] += T[q];   // Q_a = T + t_a
    JF_UNROLL for (int a = 0; a < 4; a++) JF_UNROLL for (int b = a + 1; b < 4; b++) {
        const double (&qa)[6] = t[a];
        const double (&qb)[6] = t[b];
        double o[3];
        o[0] = qb[0] * h[a][0] + qb[3] * h[a][1] + qb[5] * h[a][2] + qa[0] * h[b][0] + qa[3] * h[b][1] + qa[5] * h[b][2];
        o[1] = qb[3] * h[a][0] + qb[1] * h[a][1] + qb[4] * h[a][2] + qa[3] * h[b][0] + qa[1] * h[b][1] + qa[4] * h[b][2];
        o[2] = qb[5] * h[a][0] + qb[4] * h[a][1] + qb[2] * h[a][2] + qa[5] * h[b][0] + qa[4] * h[b][1] + qa[2] * h[b][2];
        out(t10_edge(a, b), o[0], o[1], o[2]);
    }
}

// ------------------------------------------------------------------------------------------------
// Tet4, GLTET1 (src/quadrature/gltet.jl:7-11): constant gradient.
// ------------------------------------------------------------------------------------------------
template <class Pt, class FLD, class OUT>
JF_HD bool tet4_general(const Pt &pt0, long long elem, const FLD (&F)[Pt::NF], const FLD &X, OUT &&out) {
    constexpr int NF = Pt::NF;
    Pt pt = pt0;
    pt.load(elem);
    double J[3][3], iJ[3][3];
    JF_UNROLL for (int a = 0; a < 3; a++) JF_UNROLL for (int c = 0; c < 3; c++) J[a][c] = X(a + 1, c) - X(0, c);
    double det = inv3x3(J, iJ);
    double G[NF][3][3];
    JF_UNROLL for (int f = 0; f < NF; f++) {
        double Hh[3][3];
        JF_UNROLL for (int a = 0; a < 3; a++) JF_UNROLL for (int c = 0; c < 3; c++) Hh[a][c] = F[f](a + 1, c) - F[f](0, c);
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++)
            G[f][i][j] = iJ[j][0] * Hh[0][i] + iJ[j][1] * Hh[1][i] + iJ[j][2] * Hh[2][i];
    }
    double P[3][3];
    bool ok = pt.eval(elem, G, P);
    double w = det * (1.0 / 6.0);
    double t[4][3];
    JF_UNROLL for (int i = 0; i < 3; i++) {
        t[1][i] = w * (P[i][0] * iJ[0][0] + P[i][1] * iJ[1][0] + P[i][2] * iJ[2][0]);
        t[2][i] = w * (P[i][0] * iJ[0][1] + P[i][1] * iJ[1][1] + P[i][2] * iJ[2][1]);
        t[3][i] = w * (P[i][0] * iJ[0][2] + P[i][1] * iJ[1][2] + P[i][2] * iJ[2][2]);
        t[0][i] = -(t[1][i] + t[2][i] + t[3][i]);
    }
    JF_UNROLL for (int k = 0; k < 4; k++) out(k, t[k][0], t[k][1], t[k][2]);
    return ok;
}

// ------------------------------------------------------------------------------------------------
// Hex8, trilinear isoparametric, GLHEX8 (+-1/sqrt(3), weight 1; src/quadrature/quaddata.jl:4-5).
// Reference-space derivatives through the modal (1,u,v,w,uv,uw,vw,uvw) coefficients of each field:
// d/du at (.,v,w) = c1 + c4 v + c5 w + c7 v w, so the 8 Gauss-point gradients of a field cost ~60 flops
// instead of 8 x 8 x 3.  Node order (-,-,-),(+,-,-),(+,+,-),(-,+,-),(-,-,+),... (lagrange_generated.jl:289-290).
// ------------------------------------------------------------------------------------------------
JF_HD void hex8_modal(const double (&q)[8], double (&m)[8]) {
    // m = 1/8 * sum_i sign_i * q_i for monomials 1,u,v,w,uv,uw,vw,uvw
    double a0 = q[0] + q[1], a1 = q[1] - q[0], a2 = q[3] + q[2], a3 = q[2] - q[3];
    double a4 = q[4] + q[5], a5 = q[5] - q[4], a6 = q[7] + q[6], a7 = q[6] - q[7];
    // v-direction: (a0: v=-1 , a2: v=+1) etc.
    double b0 = a0 + a2, b1 = a1 + a3, b2 = a2 - a0, b3 = a3 - a1;
    double b4 = a4 + a6, b5 = a5 + a7, b6 = a6 - a4, b7 = a7 - a5;
    m[0] = 0.125 * (b0 + b4); m[1] = 0.125 * (b1 + b5); m[2] = 0.125 * (b2 + b6); m[4] = 0.125 * (b3 + b7);
    m[3] = 0.125 * (b4 - b0); m[5] = 0.125 * (b5 - b1); m[6] = 0.125 * (b6 - b2); m[7] = 0.125 * (b7 - b3);
}

#define HEX_GA 0.5773502691896258

template <class Pt, class FLD, class OUT>
JF_HD bool hex8_general(const Pt &pt0, long long elem, const FLD (&F)[Pt::NF], const FLD &X, OUT &&out) {
    constexpr int NF = Pt::NF;
    Pt pt = pt0;
    pt.load(elem);
    double mx[3][8], mu_[NF][3][8];
    JF_UNROLL for (int c = 0; c < 3; c++) {
        double q[8];
        JF_UNROLL for (int k = 0; k < 8; k++) q[k] = X(k, c);
        hex8_modal(q, mx[c]);
        JF_UNROLL for (int f = 0; f < NF; f++) {
            JF_UNROLL for (int k = 0; k < 8; k++) q[k] = F[f](k, c);
            hex8_modal(q, mu_[f][c]);
        }
    }
    // modal accumulators of the transposed operation: R[i][a][m] collects sum_g P_g[i][a] * (dmonomial/dxi_a)(xi_g)
    double Ru[3][4], Rv[3][4], Rw[3][4];   // for d/du: coefficients of (1, v, w, vw); d/dv: (1,u,w,uw); d/dw: (1,u,v,uv)
    JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int m = 0; m < 4; m++) { Ru[i][m] = 0; Rv[i][m] = 0; Rw[i][m] = 0; }
    bool ok = true;
    JF_UNROLL for (int g = 0; g < 8; g++) {   // first index fastest (glquad.jl:12-15)
        const double u = (g & 1) ? HEX_GA : -HEX_GA, v = (g & 2) ? HEX_GA : -HEX_GA, w = (g & 4) ? HEX_GA : -HEX_GA;
        double J[3][3], iJ[3][3];
        JF_UNROLL for (int c = 0; c < 3; c++) {
            J[0][c] = mx[c][1] + mx[c][4] * v + mx[c][5] * w + mx[c][7] * (v * w);
            J[1][c] = mx[c][2] + mx[c][4] * u + mx[c][6] * w + mx[c][7] * (u * w);
            J[2][c] = mx[c][3] + mx[c][5] * u + mx[c][6] * v + mx[c][7] * (u * v);
        }
        double det = inv3x3(J, iJ);
        double G[NF][3][3];
        JF_UNROLL for (int f = 0; f < NF; f++) {
            double Hh[3][3];
            JF_UNROLL for (int c = 0; c < 3; c++) {
                Hh[0][c] = mu_[f][c][1] + mu_[f][c][4] * v + mu_[f][c][5] * w + mu_[f][c][7] * (v * w);
                Hh[1][c] = mu_[f][c][2] + mu_[f][c][4] * u + mu_[f][c][6] * w + mu_[f][c][7] * (u * w);
                Hh[2][c] = mu_[f][c][3] + mu_[f][c][5] * u + mu_[f][c][6] * v + mu_[f][c][7] * (u * v);
            }
            JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++)
                G[f][i][j] = iJ[j][0] * Hh[0][i] + iJ[j][1] * Hh[1][i] + iJ[j][2] * Hh[2][i];
        }
        double P[3][3];
        ok &= pt.eval(elem * 8 + g, G, P);
        JF_UNROLL for (int i = 0; i < 3; i++) {
            double t0 = det * (P[i][0] * iJ[0][0] + P[i][1] * iJ[1][0] + P[i][2] * iJ[2][0]);   // coefficient of dN/du
            double t1 = det * (P[i][0] * iJ[0][1] + P[i][1] * iJ[1][1] + P[i][2] * iJ[2][1]);   // dN/dv
            double t2 = det * (P[i][0] * iJ[0][2] + P[i][1] * iJ[1][2] + P[i][2] * iJ[2][2]);   // dN/dw
            Ru[i][0] += t0; Ru[i][1] += t0 * v; Ru[i][2] += t0 * w; Ru[i][3] += t0 * (v * w);
            Rv[i][0] += t1; Rv[i][1] += t1 * u; Rv[i][2] += t1 * w; Rv[i][3] += t1 * (u * w);
            Rw[i][0] += t2; Rw[i][1] += t2 * u; Rw[i][2] += t2 * v; Rw[i][3] += t2 * (u * v);
        }
    }
    // f_k[i] = sum over monomials: N_k = 1/8 (1 + a u)(1 + b v)(1 + c w) with node signs (a,b,c):
    // dN_k/du = 1/8 a (1 + b v + c w + b c v w)  -> contributes a/8 (Ru0 + b Ru1 + c Ru2 + b c Ru3), same for v, w.
    JF_UNROLL for (int k = 0; k < 8; k++) {
        const double a = (k == 1 || k == 2 || k == 5 || k == 6) ? 1.0 : -1.0;
        const double b = (k == 2 || k == 3 || k == 6 || k == 7) ? 1.0 : -1.0;
        const double c = (k >= 4) ? 1.0 : -1.0;
        double r[3];
        JF_UNROLL for (int i = 0; i < 3; i++)
            r[i] = 0.125 * (a * (Ru[i][0] + b * Ru[i][1] + c * Ru[i][2] + (b * c) * Ru[i][3])
                          + b * (Rv[i][0] + a * Rv[i][1] + c * Rv[i][2] + (a * c) * Rv[i][3])
                          + c * (Rw[i][0] + a * Rw[i][1] + b * Rw[i][2] + (a * b) * Rw[i][3]));
        out(k, r[0], r[1], r[2]);
    }
    return ok;
}

// ------------------------------------------------------------------------------------------------
// Hex8 parallelepiped (constant Jacobian) + linear elasticity.  With constant inv(J) the displacement gradient is a
// polynomial in the 7 monomials (1,u,v,w,uv,uw,vw) whose coefficients come straight from the modal coefficients of u,
// the stress is linear in it, and the 2x2x2 Gauss sum of (coefficient x test monomial) is diagonal:
//   sum_g phi_m phi_n = 8 gamma^(2 deg m) delta_mn   (gamma^2 = 1/3)
// so the 8 Gauss-point loops of hex8_general collapse to 7 stress evaluations and 12 small mat-vecs
// (~800 fp64 instructions per element instead of ~2200).  X(0..3) = coordinates of element nodes 0, 1, 3, 4.
// ------------------------------------------------------------------------------------------------
template <class FLD, class XFLD, class OUT>
JF_HD void hex8_affine_linear(double la, double mu, const FLD &U, const XFLD &X, OUT &&out) {
    double J[3][3], iJ[3][3];
    JF_UNROLL for (int a = 0; a < 3; a++) JF_UNROLL for (int c = 0; c < 3; c++) J[a][c] = 0.5 * (X(a + 1, c) - X(0, c));   // J[a][b] = dx_b/dxi_a
    const double det = inv3x3(J, iJ);
    double m[3][8];
    JF_UNROLL for (int c = 0; c < 3; c++) {
        double q[8];
        JF_UNROLL for (int k = 0; k < 8; k++) q[k] = U(k, c);
        hex8_modal(q, m[c]);
    }
    // t_a[i] coefficient of monomial mu, already times det * sum_g phi_mu^2:  R = sc * det * P(G) . iJ[:,a]
    auto stress_cols = [&](const double (&G)[3][3], double sc, double (&T)[3][3]) {   // T[i][a]
        const double las = la * sc * det, mus = mu * sc * det;
        const double tr = las * (G[0][0] + G[1][1] + G[2][2]);
        double P[3][3];
        P[0][0] = tr + 2 * mus * G[0][0]; P[1][1] = tr + 2 * mus * G[1][1]; P[2][2] = tr + 2 * mus * G[2][2];
        P[0][1] = P[1][0] = mus * (G[0][1] + G[1][0]); P[1][2] = P[2][1] = mus * (G[1][2] + G[2][1]); P[0][2] = P[2][0] = mus * (G[0][2] + G[2][0]);
        JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int a = 0; a < 3; a++)
            T[i][a] = P[i][0] * iJ[0][a] + P[i][1] * iJ[1][a] + P[i][2] * iJ[2][a];
    };
    double Ru[3][4], Rv[3][4], Rw[3][4];   // as in hex8_general: d/du: (1, v, w, vw); d/dv: (1, u, w, uw); d/dw: (1, u, v, uv)
    double G[3][3], T[3][3];
    // monomial 1
    JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) G[i][j] = iJ[j][0] * m[i][1] + iJ[j][1] * m[i][2] + iJ[j][2] * m[i][3];
    stress_cols(G, 8.0, T);
    JF_UNROLL for (int i = 0; i < 3; i++) { Ru[i][0] = T[i][0]; Rv[i][0] = T[i][1]; Rw[i][0] = T[i][2]; }
    // monomial u: d/dv and d/dw see it
    JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) G[i][j] = iJ[j][1] * m[i][4] + iJ[j][2] * m[i][5];
    stress_cols(G, 8.0 / 3.0, T);
    JF_UNROLL for (int i = 0; i < 3; i++) { Rv[i][1] = T[i][1]; Rw[i][1] = T[i][2]; }
    // monomial v
    JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) G[i][j] = iJ[j][0] * m[i][4] + iJ[j][2] * m[i][6];
    stress_cols(G, 8.0 / 3.0, T);
    JF_UNROLL for (int i = 0; i < 3; i++) { Ru[i][1] = T[i][0]; Rw[i][2] = T[i][2]; }
    // monomial w
    JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) G[i][j] = iJ[j][0] * m[i][5] + iJ[j][1] * m[i][6];
    stress_cols(G, 8.0 / 3.0, T);
    JF_UNROLL for (int i = 0; i < 3; i++) { Ru[i][2] = T[i][0]; Rv[i][2] = T[i][1]; }
    // monomials uv, uw, vw (only m7)
    JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) G[i][j] = iJ[j][2] * m[i][7];
    stress_cols(G, 8.0 / 9.0, T);
    JF_UNROLL for (int i = 0; i < 3; i++) Rw[i][3] = T[i][2];
    JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) G[i][j] = iJ[j][1] * m[i][7];
    stress_cols(G, 8.0 / 9.0, T);
    JF_UNROLL for (int i = 0; i < 3; i++) Rv[i][3] = T[i][1];
    JF_UNROLL for (int i = 0; i < 3; i++) JF_UNROLL for (int j = 0; j < 3; j++) G[i][j] = iJ[j][0] * m[i][7];
    stress_cols(G, 8.0 / 9.0, T);
    JF_UNROLL for (int i = 0; i < 3; i++) Ru[i][3] = T[i][0];
    JF_UNROLL for (int k = 0; k < 8; k++) {
        const double a = (k == 1 || k == 2 || k == 5 || k == 6) ? 1.0 : -1.0;
        const double b = (k == 2 || k == 3 || k == 6 || k == 7) ? 1.0 : -1.0;
        const double c = (k >= 4) ? 1.0 : -1.0;
        double r[3];
        JF_UNROLL for (int i = 0; i < 3; i++)
            r[i] = 0.125 * (a * (Ru[i][0] + b * Ru[i][1] + c * Ru[i][2] + (b * c) * Ru[i][3])
                          + b * (Rv[i][0] + a * Rv[i][1] + c * Rv[i][2] + (a * c) * Rv[i][3])
                          + c * (Rw[i][0] + a * Rw[i][1] + b * Rw[i][2] + (a * b) * Rw[i][3]));
        out(k, r[0], r[1], r[2]);
    }
}

}  // namespace jf

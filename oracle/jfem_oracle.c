/*
 * jfem_oracle.c -- CPU ORACLE for the JuliaFEM 3D-elasticity hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product library
 * (libjfem_b200.so) never links or calls anything in this directory.
 *
 * It is a plain-C restatement (not a copy; the reference is Julia) of the arithmetic in
 * the reference, each function citing the reference file:line it follows (paths relative
 * to the JuliaFEM.jl checkout):
 *
 *   shape functions      src/basis/lagrange_generated.jl:239-254 (Tet4) :267-282 (Tet10) :295-310 (Hex8)
 *   quadrature           src/quadrature/gltet.jl:7-25, src/quadrature/glquad.jl:6-44, src/quadrature/quaddata.jl:4-5
 *   rule selection       src/elements/integrate.jl:19,26-27 (Hex8->GLHEX8, Tet4->GLTET1, Tet10->GLTET4)
 *   Jacobian / gradients src/basis/math.jl:47-54 (J = sum dN_i (x) X_i), :185-212 (grad = inv(J).dN), :249-255
 *   element integration  src/problems_elasticity.jl:203-451 (Voigt BL, Km += w BL' D BL, f_int += w BL' s)
 *   block-form check     src/physics/assembly_helpers.jl:201-226, :263-283
 *   materials            src/materials/linear_elastic.jl:82,97,136-158
 *                        src/materials/neo_hookean.jl:129-143,205-231 (AD replaced by closed form, checked against forward-mode AD and FD)
 *                        src/materials/perfect_plasticity.jl:247-357
 *   scatter / pattern    src/sparse/sparse.jl:53-55,121-132,178-181 ; src/assembly/problems.jl:466-478
 *   symmetrisation       src/solvers.jl:289-292
 *   Dirichlet + CG       ext/JuliaFEMCUDAExt.jl:423-435,531-577 ; src/backend/cpu.jl:221-254
 *
 * PARITY PIN STATUS: the reference cannot run here (no Julia; the package does not load as shipped).  Pinned against values
 * the reference itself holds (tests/golden/pins.json, tests/test_oracle_pins.py):
 *   - END TO END: examples/linear_static.jl:133, max |u| = 2.4052929896922337 on JuliaFEMSMP18.med (Tet10 shape functions and
 *     node order, GLTET4, D, scatter + dof numbering, consistent body load, Dirichlet elimination, solve) -- reproduced by this
 *     oracle to 6e-12 relative and by the GPU path to 1e-8 (tests/test_gpu_host_api.py);
 *   - LE uniaxial / shear / tangent identity, PP return-map values + on-surface identity, the Tet10 consistent-mass table of
 *     src/assembly/assembly.jl:139-149, quadrature constants, the reference's tet10.inp fixture mesh;
 *   - Neo-Hookean: the closed form below equals second-order forward-mode AD of the reference's strain_energy (what
 *     Tensors.hessian computes, neo_hookean.jl:222) to 1e-12.
 * Not pinned by a reference-held value (none survives in the snapshot): individual Ke entries, CSR values and K.u of other
 * meshes; those rest on this restatement plus analytic identities (symmetry, rigid modes, patch test, block form == Voigt form).
 *
 * Layout conventions: node coordinates X[i*3+b]; element connectivity conn[e*nnpe+k], 0-based
 * here (the Python wrapper converts from the reference's 1-based ids); dof = 3*node + c
 * (0-based form of 3*(node-1)+c, src/assembly/problems.jl:476); Ke column-major ndof x ndof as
 * Julia stores it; Voigt order 11,22,33,12,23,13 with engineering shears
 * (src/problems_elasticity.jl:189-197).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_TET4 4
#define ORC_HEX8 8
#define ORC_TET10 10

#define ORC_MAT_LE 0
#define ORC_MAT_NH 1
#define ORC_MAT_PP 2

#define ORC_MAXN 10
#define ORC_MAXDOF 30
#define ORC_MAXGP 8
#define ORC_NSTATE 13 /* eps_p(6 tensor comps 11,22,33,12,23,13) alpha(6) kappa(1) */

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ shape functions */

/* src/basis/lagrange_generated.jl:267-282 (values), expanded monomials as generated */
void orc_shape_N(int et, const double xi[3], double *N) {
    double u = xi[0], v = xi[1], w = xi[2];
    if (et == ORC_TET4) {
        N[0] = 1 + -1.0 * u + -1.0 * v + -1.0 * w; N[1] = u; N[2] = v; N[3] = w;
    } else if (et == ORC_TET10) {
        N[0] = 1 + -3.0 * u + -3.0 * v + -3.0 * w + 2.0 * u * u + 2.0 * v * v + 2.0 * w * w + 4.0 * (u * v) + 4.0 * (u * w) + 4.0 * (v * w);
        N[1] = -1.0 * u + 2.0 * u * u;
        N[2] = -1.0 * v + 2.0 * v * v;
        N[3] = -1.0 * w + 2.0 * w * w;
        N[4] = 4.0 * u + -4.0 * u * u + -4.0 * (u * v) + -4.0 * (u * w);
        N[5] = 4.0 * (u * v);
        N[6] = 4.0 * v + -4.0 * v * v + -4.0 * (u * v) + -4.0 * (v * w);
        N[7] = 4.0 * w + -4.0 * w * w + -4.0 * (u * w) + -4.0 * (v * w);
        N[8] = 4.0 * (u * w);
        N[9] = 4.0 * (v * w);
    } else { /* Hex8, :295-302 */
        static const double s[8][3] = {{-1,-1,-1},{1,-1,-1},{1,1,-1},{-1,1,-1},{-1,-1,1},{1,-1,1},{1,1,1},{-1,1,1}};
        for (int i = 0; i < 8; i++)
            N[i] = 0.125 + 0.125 * s[i][0] * u + 0.125 * s[i][1] * v + 0.125 * s[i][2] * w
                 + 0.125 * s[i][0] * s[i][1] * (u * v) + 0.125 * s[i][0] * s[i][2] * (u * w)
                 + 0.125 * s[i][1] * s[i][2] * (v * w) + 0.125 * s[i][0] * s[i][1] * s[i][2] * (u * v * w);
    }
}

/* dN[i*3+a] = dN_i/dxi_a.  src/basis/lagrange_generated.jl:247-254, :275-282, :303-310 */
void orc_shape_dN(int et, const double xi[3], double *dN) {
    double u = xi[0], v = xi[1], w = xi[2];
    if (et == ORC_TET4) {
        static const double d[12] = {-1,-1,-1, 1,0,0, 0,1,0, 0,0,1};
        memcpy(dN, d, sizeof d);
    } else if (et == ORC_TET10) {
        double d[30] = {
            -3.0 + 2.0 * (2 * u) + 4.0 * v + 4.0 * w, -3.0 + 2.0 * (2 * v) + 4.0 * u + 4.0 * w, -3.0 + 2.0 * (2 * w) + 4.0 * u + 4.0 * v,
            -1.0 + 2.0 * (2 * u), 0, 0,
            0, -1.0 + 2.0 * (2 * v), 0,
            0, 0, -1.0 + 2.0 * (2 * w),
            4.0 + -4.0 * (2 * u) + -4.0 * v + -4.0 * w, -4.0 * u, -4.0 * u,
            4.0 * v, 4.0 * u, 0,
            -4.0 * v, 4.0 + -4.0 * (2 * v) + -4.0 * u + -4.0 * w, -4.0 * v,
            -4.0 * w, -4.0 * w, 4.0 + -4.0 * (2 * w) + -4.0 * u + -4.0 * v,
            4.0 * w, 0, 4.0 * u,
            0, 4.0 * w, 4.0 * v};
        memcpy(dN, d, sizeof d);
    } else {
        static const double s[8][3] = {{-1,-1,-1},{1,-1,-1},{1,1,-1},{-1,1,-1},{-1,-1,1},{1,-1,1},{1,1,1},{-1,1,1}};
        for (int i = 0; i < 8; i++) {
            double a = s[i][0], b = s[i][1], c = s[i][2];
            dN[i * 3 + 0] = 0.125 * a + 0.125 * a * b * v + 0.125 * a * c * w + 0.125 * a * b * c * (v * w);
            dN[i * 3 + 1] = 0.125 * b + 0.125 * a * b * u + 0.125 * b * c * w + 0.125 * a * b * c * (u * w);
            dN[i * 3 + 2] = 0.125 * c + 0.125 * a * c * u + 0.125 * b * c * v + 0.125 * a * b * c * (u * v);
        }
    }
}

/* Default rule per element type (src/elements/integrate.jl:19,26-27).  Returns #points.
 * GLTET4 constants from the sqrt expressions of src/quadrature/gltet.jl:19-20;
 * GLHEX8 first index fastest (src/quadrature/glquad.jl:12-15, CartesianIndices). */
int orc_quadrature(int et, double *w, double *xi /* ngp x 3 */) {
    if (et == ORC_TET4) {
        w[0] = 1.0 / 6.0; xi[0] = xi[1] = xi[2] = 1.0 / 4.0; return 1;
    }
    if (et == ORC_TET10) {
        double a = (5.0 + 3.0 * sqrt(5.0)) / 20.0, b = (5.0 - sqrt(5.0)) / 20.0;
        double p[12] = {a, b, b, b, a, b, b, b, a, b, b, b};
        memcpy(xi, p, sizeof p);
        for (int i = 0; i < 4; i++) w[i] = 1.0 / 24.0;
        return 4;
    }
    const double g[2] = {-0.5773502691896258, 0.5773502691896258};
    int q = 0;
    for (int k = 0; k < 2; k++) for (int j = 0; j < 2; j++) for (int i = 0; i < 2; i++) {
        xi[q * 3 + 0] = g[i]; xi[q * 3 + 1] = g[j]; xi[q * 3 + 2] = g[k]; w[q] = 1.0; q++;
    }
    return 8;
}

/* ------------------------------------------------------------------ small tensor helpers */

static double det3(const double A[9]) {
    return A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) + A[2] * (A[3] * A[7] - A[4] * A[6]);
}
static void inv3(const double A[9], double B[9], double *det) {
    double d = det3(A), r = 1.0 / d;
    B[0] = (A[4] * A[8] - A[5] * A[7]) * r; B[1] = (A[2] * A[7] - A[1] * A[8]) * r; B[2] = (A[1] * A[5] - A[2] * A[4]) * r;
    B[3] = (A[5] * A[6] - A[3] * A[8]) * r; B[4] = (A[0] * A[8] - A[2] * A[6]) * r; B[5] = (A[2] * A[3] - A[0] * A[5]) * r;
    B[6] = (A[3] * A[7] - A[4] * A[6]) * r; B[7] = (A[1] * A[6] - A[0] * A[7]) * r; B[8] = (A[0] * A[4] - A[1] * A[3]) * r;
    *det = d;
}
/* Voigt index of tensor component (i,j): order 11,22,33,12,23,13 */
static const int VG[3][3] = {{0, 3, 5}, {3, 1, 4}, {5, 4, 2}};

/* ------------------------------------------------------------------ materials */
/* All three return the stress as tensor components s[6] (11,22,33,12,23,13) and the tangent
 * as a 6x6 Voigt matrix D[I*6+J] = DD_ijkl, to be applied to engineering-shear strain vectors. */

void orc_lame(double E, double nu, double *la, double *mu) {
    *la = E * nu / ((1.0 + nu) * (1.0 - 2.0 * nu)); /* src/materials/linear_elastic.jl:82; src/problems_elasticity.jl:313 */
    *mu = E / (2.0 * (1.0 + nu));                   /* :97 ; :314 */
}

static void elastic_D(double la, double mu, double D[36]) {
    memset(D, 0, 36 * sizeof(double));              /* src/problems_elasticity.jl:315-317 */
    D[0] = D[7] = D[14] = 2 * mu + la;
    D[21] = D[28] = D[35] = mu;
    D[1] = D[6] = D[8] = D[13] = D[2] = D[12] = la;
}

/* src/materials/linear_elastic.jl:136-158.  eps = tensor components (not engineering). */
void orc_le_stress(double E, double nu, const double eps[6], double s[6], double D[36]) {
    double la, mu; orc_lame(E, nu, &la, &mu);
    double tr = eps[0] + eps[1] + eps[2];
    for (int i = 0; i < 3; i++) s[i] = la * tr + 2 * mu * eps[i];
    for (int i = 3; i < 6; i++) s[i] = 2 * mu * eps[i];
    elastic_D(la, mu, D);
}

/* psi(C), src/materials/neo_hookean.jl:129-143.  C as symmetric 6-vector. Returns NaN if J<=0 */
double orc_nh_energy(double mu, double la, const double C[6]) {
    double M[9] = {C[0], C[3], C[5], C[3], C[1], C[4], C[5], C[4], C[2]};
    double d = det3(M);
    if (!(d > 0.0)) return NAN;
    double J = sqrt(d), lj = log(J);
    return mu / 2 * (C[0] + C[1] + C[2] - 3) - mu * lj + la / 2 * lj * lj;
}

/* src/materials/neo_hookean.jl:205-231: S = 2 dpsi/dC, DD = 4 d2psi/dC2 at C = 2E + I.
 * The reference obtains both by Tensors.hessian (forward-mode AD, Tensors.jl 1.16.2, not in the
 * tree).  Closed form used here: S = mu (I - C^-1) + la lnJ C^-1,
 * DD_ijkl = la Ci_ij Ci_kl + (mu - la lnJ)(Ci_ik Ci_jl + Ci_il Ci_jk); tests check it against
 * central differences of orc_nh_energy.  Returns 1 if J <= 0 (reference throws DomainError :137). */
int orc_nh_stress(double mu, double la, const double E[6], double S[6], double D[36]) {
    double C[9] = {2 * E[0] + 1, 2 * E[3], 2 * E[5], 2 * E[3], 2 * E[1] + 1, 2 * E[4], 2 * E[5], 2 * E[4], 2 * E[2] + 1};
    double Ci[9], d;
    inv3(C, Ci, &d);
    if (!(d > 0.0)) return 1;
    double lj = 0.5 * log(d);
    for (int i = 0; i < 3; i++) for (int j = i; j < 3; j++)
        S[VG[i][j]] = mu * ((i == j ? 1.0 : 0.0) - Ci[i * 3 + j]) + la * lj * Ci[i * 3 + j];
    double c2 = mu - la * lj;
    for (int i = 0; i < 3; i++) for (int j = i; j < 3; j++) for (int k = 0; k < 3; k++) for (int l = k; l < 3; l++)
        D[VG[i][j] * 6 + VG[k][l]] = la * Ci[i * 3 + j] * Ci[k * 3 + l]
            + c2 * (Ci[i * 3 + k] * Ci[j * 3 + l] + Ci[i * 3 + l] * Ci[j * 3 + k]);
    return 0;
}

/* src/materials/perfect_plasticity.jl:247-341.  par = {E, nu, sigma_y, H}; eps tensor comps;
 * state = eps_p[6], alpha[6], kappa.  Returns 1 if the step was plastic. */
int orc_pp_stress(const double par[4], const double eps[6], const double *st_old, double s[6], double D[36], double *st_new) {
    double la, mu; orc_lame(par[0], par[1], &la, &mu);
    double sy = par[2], H = par[3];
    double zero[ORC_NSTATE] = {0};
    if (!st_old) st_old = zero;
    const double *ep = st_old, *al = st_old + 6;
    double ee[6], st[6], sd[6];
    for (int i = 0; i < 6; i++) ee[i] = eps[i] - ep[i];                       /* :270 */
    double tr = ee[0] + ee[1] + ee[2];
    for (int i = 0; i < 3; i++) st[i] = la * tr + 2 * mu * ee[i];            /* :275 */
    for (int i = 3; i < 6; i++) st[i] = 2 * mu * ee[i];
    for (int i = 0; i < 6; i++) sd[i] = st[i] - al[i];
    double m = (sd[0] + sd[1] + sd[2]) / 3.0;                                 /* dev(), :279 */
    sd[0] -= m; sd[1] -= m; sd[2] -= m;
    double nn = sd[0] * sd[0] + sd[1] * sd[1] + sd[2] * sd[2] + 2 * (sd[3] * sd[3] + sd[4] * sd[4] + sd[5] * sd[5]);
    double q = sqrt(3.0 / 2.0) * sqrt(nn);                                     /* :282 */
    double f = q - sy;                                                         /* :285 */
    elastic_D(la, mu, D);
    if (f <= 0.0) {                                                            /* :288-295 */
        memcpy(s, st, 6 * sizeof(double));
        if (st_new) memcpy(st_new, st_old, ORC_NSTATE * sizeof(double));
        return 0;
    }
    double n[6];
    for (int i = 0; i < 6; i++) n[i] = sd[i] / q;                              /* :299 */
    double den = 2 * mu + (2.0 / 3.0) * H;
    double dl = f / den;                                                       /* :315 */
    for (int i = 0; i < 6; i++) s[i] = st[i] - 2 * mu * dl * n[i];            /* :318 */
    if (st_new) {
        for (int i = 0; i < 6; i++) st_new[6 + i] = al[i] + (2.0 / 3.0) * H * dl * n[i];  /* :321 */
        for (int i = 0; i < 6; i++) st_new[i] = ep[i] + dl * n[i];                      /* :324 */
        st_new[12] = st_old[12] + dl;                                                   /* :327 */
    }
    double c = 4 * mu * mu / den;                                              /* :337 */
    for (int I = 0; I < 6; I++) for (int J = 0; J < 6; J++) D[I * 6 + J] -= c * n[I] * n[J];
    return 1;
}

/* ------------------------------------------------------------------ element integration */

typedef struct {
    int kind;            /* ORC_MAT_* */
    double par[4];       /* LE: E,nu ; NH: E,nu (mu,la via orc_lame, neo_hookean.jl:87-100) ; PP: E,nu,sy,H */
    int finite_strain;   /* props.finite_strain  (src/problems_elasticity.jl:258) */
    int geometric;       /* props.geometric_stiffness (:378) */
} orc_material;

/* One Gauss point of src/problems_elasticity.jl:245-409.  X,u: nn x 3.  Adds into Km (col-major
 * ndof x ndof, may be NULL), Kg (may be NULL), fint (may be NULL).  st_old/st_new: 13 doubles or NULL.
 * Returns 0 ok, 1 = invalid deformation (NH J<=0). */
static int gauss_point(int et, int nn, const double *X, const double *u, const orc_material *mat,
                       double wq, const double xi[3], const double *st_old, double *st_new,
                       double *Km, double *Kg, double *fint) {
    int ndof = 3 * nn;
    double dN[ORC_MAXN * 3], J[9] = {0}, iJ[9], detJ, G[ORC_MAXN * 3];
    orc_shape_dN(et, xi, dN);
    for (int i = 0; i < nn; i++)                                  /* src/basis/math.jl:47-54 : J[a][b] += dN_i[a] X_i[b] */
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) J[a * 3 + b] += dN[i * 3 + a] * X[i * 3 + b];
    inv3(J, iJ, &detJ);                                           /* math.jl:198, :202 */
    for (int i = 0; i < nn; i++) for (int b = 0; b < 3; b++)      /* math.jl:199-201 : grad_i = invJ . dN_i */
        G[i * 3 + b] = iJ[b * 3 + 0] * dN[i * 3 + 0] + iJ[b * 3 + 1] * dN[i * 3 + 1] + iJ[b * 3 + 2] * dN[i * 3 + 2];
    double w = wq * detJ;                                         /* problems_elasticity.jl:248 */
    double gu[9] = {0};                                           /* math.jl:249-255 : gradu += u_k (x) grad_k */
    for (int k = 0; k < nn; k++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) gu[i * 3 + j] += u[k * 3 + i] * G[k * 3 + j];
    double F[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, e[9];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {     /* :259 / :262 */
        double v = 0.5 * (gu[i * 3 + j] + gu[j * 3 + i]);
        if (mat->finite_strain) { double q = 0; for (int k = 0; k < 3; k++) q += gu[k * 3 + i] * gu[k * 3 + j]; v += 0.5 * q; }
        e[i * 3 + j] = v;
    }
    if (mat->finite_strain) for (int i = 0; i < 9; i++) F[i] += gu[i];   /* :260 */
    double et6[6] = {e[0], e[4], e[8], e[1], e[5], e[2]};          /* tensor comps, Voigt order */
    double BL[6 * ORC_MAXDOF];
    memset(BL, 0, sizeof BL);
    for (int i = 0; i < nn; i++) {
        const double *g = G + i * 3;
        if (mat->finite_strain) {                                  /* :270-289 */
            for (int c = 0; c < 3; c++) {
                int col = 3 * i + c;
                BL[0 * ndof + col] = F[c * 3 + 0] * g[0];
                BL[1 * ndof + col] = F[c * 3 + 1] * g[1];
                BL[2 * ndof + col] = F[c * 3 + 2] * g[2];
                BL[3 * ndof + col] = F[c * 3 + 0] * g[1] + F[c * 3 + 1] * g[0];
                BL[4 * ndof + col] = F[c * 3 + 1] * g[2] + F[c * 3 + 2] * g[1];
                BL[5 * ndof + col] = F[c * 3 + 2] * g[0] + F[c * 3 + 0] * g[2];
            }
        } else {                                                   /* :291-301 */
            BL[0 * ndof + 3 * i + 0] = g[0]; BL[1 * ndof + 3 * i + 1] = g[1]; BL[2 * ndof + 3 * i + 2] = g[2];
            BL[3 * ndof + 3 * i + 0] = g[1]; BL[3 * ndof + 3 * i + 1] = g[0];
            BL[4 * ndof + 3 * i + 1] = g[2]; BL[4 * ndof + 3 * i + 2] = g[1];
            BL[5 * ndof + 3 * i + 0] = g[2]; BL[5 * ndof + 3 * i + 2] = g[0];
        }
    }
    double s[6], D[36];
    if (mat->kind == ORC_MAT_LE) orc_le_stress(mat->par[0], mat->par[1], et6, s, D);           /* :313-332 */
    else if (mat->kind == ORC_MAT_NH) {
        double la, mu; orc_lame(mat->par[0], mat->par[1], &la, &mu);
        if (orc_nh_stress(mu, la, et6, s, D)) return 1;
    } else orc_pp_stress(mat->par, et6, st_old, s, D, st_new);
    if (Km) {                                                      /* :370-375 */
        double DB[6 * ORC_MAXDOF];
        for (int I = 0; I < 6; I++) for (int c = 0; c < ndof; c++) {
            double a = 0; for (int Jv = 0; Jv < 6; Jv++) a += D[I * 6 + Jv] * BL[Jv * ndof + c];
            DB[I * ndof + c] = a;
        }
        for (int c = 0; c < ndof; c++) for (int r = 0; r < ndof; r++) {
            double a = 0; for (int I = 0; I < 6; I++) a += BL[I * ndof + r] * DB[I * ndof + c];
            Km[c * ndof + r] += w * a;
        }
    }
    if (Kg && mat->geometric) {                                    /* :378-404 : Kg[3i+a,3j+a] += w g_i.S.g_j */
        double S3[9] = {s[0], s[3], s[5], s[3], s[1], s[4], s[5], s[4], s[2]};
        for (int i = 0; i < nn; i++) for (int j = 0; j < nn; j++) {
            double a = 0;
            for (int p = 0; p < 3; p++) for (int q = 0; q < 3; q++) a += G[i * 3 + p] * S3[p * 3 + q] * G[j * 3 + q];
            for (int c = 0; c < 3; c++) Kg[(3 * j + c) * ndof + 3 * i + c] += w * a;
        }
    }
    if (fint) for (int c = 0; c < ndof; c++) {                     /* :407-409 */
        double a = 0; for (int I = 0; I < 6; I++) a += BL[I * ndof + c] * s[I];
        fint[c] += w * a;
    }
    return 0;
}

static void mat_init(orc_material *m, int kind, const double *par, int fs, int geo) {
    m->kind = kind; memcpy(m->par, par, 4 * sizeof(double)); m->finite_strain = fs; m->geometric = geo;
}

/* Full element: Km, Kg (both ndof x ndof col-major; zeroed here) and fint.  st_old/st_new: ngp x 13. */
int orc_element(int et, const double *X, const double *u, int kind, const double *par, int finite_strain, int geometric,
                const double *st_old, double *st_new, double *Km, double *Kg, double *fint) {
    int nn = et, ndof = 3 * nn;
    orc_material m; mat_init(&m, kind, par, finite_strain, geometric);
    double w[ORC_MAXGP], xi[ORC_MAXGP * 3];
    int ng = orc_quadrature(et, w, xi), rc = 0;
    if (Km) memset(Km, 0, sizeof(double) * ndof * ndof);
    if (Kg) memset(Kg, 0, sizeof(double) * ndof * ndof);
    if (fint) memset(fint, 0, sizeof(double) * ndof);
    for (int g = 0; g < ng; g++)
        rc |= gauss_point(et, nn, X, u, &m, w[g], xi + 3 * g, st_old ? st_old + ORC_NSTATE * g : NULL,
                          st_new ? st_new + ORC_NSTATE * g : NULL, Km, Kg, fint);
    return rc;
}

/* Block form of src/physics/assembly_helpers.jl:201-226 for a single Gauss point set: used by the
 * tests as an independent cross-check of the Voigt form (linear elastic, small strain only). */
void orc_element_block_form(int et, const double *X, double E, double nu, double *Ke) {
    int nn = et, ndof = 3 * nn;
    double la, mu; orc_lame(E, nu, &la, &mu);
    double w[ORC_MAXGP], xi[ORC_MAXGP * 3];
    int ng = orc_quadrature(et, w, xi);
    memset(Ke, 0, sizeof(double) * ndof * ndof);
    for (int g = 0; g < ng; g++) {
        double dN[ORC_MAXN * 3], J[9] = {0}, iJ[9], detJ, G[ORC_MAXN * 3];
        orc_shape_dN(et, xi + 3 * g, dN);
        for (int i = 0; i < nn; i++) for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) J[a * 3 + b] += dN[i * 3 + a] * X[i * 3 + b];
        inv3(J, iJ, &detJ);
        for (int i = 0; i < nn; i++) for (int b = 0; b < 3; b++)
            G[i * 3 + b] = iJ[b * 3 + 0] * dN[i * 3 + 0] + iJ[b * 3 + 1] * dN[i * 3 + 1] + iJ[b * 3 + 2] * dN[i * 3 + 2];
        for (int i = 0; i < nn; i++) for (int j = 0; j < nn; j++) for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
            double acc = 0;                                        /* assembly_helpers.jl:218 : G_i[k] C[a,k,b,l] G_j[l] */
            for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) {
                double C = la * (a == k) * (b == l) + mu * ((a == b) * (k == l) + (a == l) * (k == b));
                acc += G[i * 3 + k] * C * G[j * 3 + l];
            }
            Ke[(3 * j + b) * ndof + 3 * i + a] += w[g] * detJ * acc;
        }
    }
}

/* ------------------------------------------------------------------ global pattern / assembly */

/* Node adjacency (node -> sorted unique coupled nodes), from which the dof pattern follows.
 * Pattern = union over elements of gdofs x gdofs with every entry kept, zeros included
 * (src/sparse/sparse.jl:121-132 pushes all ndofs^2 triplets; sparse() at :53-55 keeps stored zeros).
 * CSC of a structurally symmetric matrix == CSR.  Call with rowptr!=NULL, colind==NULL to count. */
static int cmp_i64(const void *a, const void *b) { int64_t x = *(const int64_t *)a, y = *(const int64_t *)b; return (x > y) - (x < y); }

int orc_csr_pattern(int et, int64_t n_nodes, int64_t n_elems, const int32_t *conn, int64_t *rowptr, int32_t *colind) {
    int nn = et;
    int64_t *cnt = calloc(n_nodes + 1, sizeof(int64_t));
    for (int64_t e = 0; e < n_elems; e++) for (int k = 0; k < nn; k++) cnt[conn[e * nn + k] + 1] += nn;
    for (int64_t i = 0; i < n_nodes; i++) cnt[i + 1] += cnt[i];
    int64_t *pairs = malloc(sizeof(int64_t) * (cnt[n_nodes] ? cnt[n_nodes] : 1));
    int64_t *fill = malloc(sizeof(int64_t) * (n_nodes + 1));
    memcpy(fill, cnt, sizeof(int64_t) * (n_nodes + 1));
    for (int64_t e = 0; e < n_elems; e++) for (int k = 0; k < nn; k++) {
        int32_t a = conn[e * nn + k];
        for (int l = 0; l < nn; l++) pairs[fill[a]++] = conn[e * nn + l];
    }
    int64_t *nadj = calloc(n_nodes + 1, sizeof(int64_t));
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t a = 0; a < n_nodes; a++) {
        int64_t lo = cnt[a], hi = cnt[a + 1], m = 0;
        qsort(pairs + lo, hi - lo, sizeof(int64_t), cmp_i64);
        for (int64_t p = lo; p < hi; p++) if (p == lo || pairs[p] != pairs[p - 1]) pairs[lo + m++] = pairs[p];
        nadj[a + 1] = m;
    }
    for (int64_t a = 0; a < n_nodes; a++) nadj[a + 1] += nadj[a];
    for (int64_t a = 0; a < n_nodes; a++) {
        int64_t deg = nadj[a + 1] - nadj[a];
        for (int c = 0; c < 3; c++) rowptr[3 * a + c] = 9 * nadj[a] + c * 3 * deg;
    }
    rowptr[3 * n_nodes] = 9 * nadj[n_nodes];
    if (colind) {
#pragma omp parallel for schedule(static)
        for (int64_t a = 0; a < n_nodes; a++) {
            int64_t deg = nadj[a + 1] - nadj[a];
            for (int c = 0; c < 3; c++) {
                int32_t *dst = colind + rowptr[3 * a + c];
                for (int64_t p = 0; p < deg; p++) for (int d = 0; d < 3; d++) dst[3 * p + d] = (int32_t)(3 * pairs[cnt[a] + p] + d);
            }
        }
    }
    free(cnt); free(pairs); free(fill); free(nadj);
    return 0;
}

static int64_t find_col(const int32_t *colind, int64_t lo, int64_t hi, int32_t c) {
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (colind[m] < c) lo = m + 1; else hi = m; }
    return lo;
}

/* Assemble K (CSR values on the pattern above) and f_int.  Element order = input order; duplicates
 * summed in element order as sparse(I,J,V) does.  K = Km (+Kg if geometric).  state: n_elems x ngp x 13.
 * symmetrise: K <- (K+K')/2 (src/solvers.jl:289-292).  Serial accumulation keeps summation order fixed. */
int orc_assemble_csr(int et, int64_t n_nodes, int64_t n_elems, const double *X, const int32_t *conn, const double *u,
                     int kind, const double *par, int finite_strain, int geometric, const double *st_old, double *st_new,
                     const int64_t *rowptr, const int32_t *colind, double *vals, double *fint, int symmetrise) {
    int nn = et, ndof = 3 * nn, rc = 0;
    double w[ORC_MAXGP], xi[ORC_MAXGP * 3];
    int ng = orc_quadrature(et, w, xi);
    int64_t nnz = rowptr[3 * n_nodes];
    if (vals) memset(vals, 0, sizeof(double) * nnz);
    if (fint) memset(fint, 0, sizeof(double) * 3 * n_nodes);
    int64_t chunk = 4096;
    double *Kbuf = malloc(sizeof(double) * chunk * ndof * ndof), *Gbuf = malloc(sizeof(double) * chunk * ndof * ndof), *fbuf = malloc(sizeof(double) * chunk * ndof);
    for (int64_t e0 = 0; e0 < n_elems; e0 += chunk) {
        int64_t e1 = e0 + chunk < n_elems ? e0 + chunk : n_elems;
#pragma omp parallel for schedule(static) reduction(| : rc)
        for (int64_t e = e0; e < e1; e++) {
            double Xe[ORC_MAXN * 3], ue[ORC_MAXN * 3];
            for (int k = 0; k < nn; k++) for (int c = 0; c < 3; c++) {
                Xe[k * 3 + c] = X[(int64_t)conn[e * nn + k] * 3 + c];
                ue[k * 3 + c] = u ? u[(int64_t)conn[e * nn + k] * 3 + c] : 0.0;
            }
            rc |= orc_element(et, Xe, ue, kind, par, finite_strain, geometric, st_old ? st_old + e * ng * ORC_NSTATE : NULL,
                              st_new ? st_new + e * ng * ORC_NSTATE : NULL, Kbuf + (e - e0) * ndof * ndof,
                              Gbuf + (e - e0) * ndof * ndof, fbuf + (e - e0) * ndof);
        }
        for (int64_t e = e0; e < e1; e++) {
            const double *Ke = Kbuf + (e - e0) * ndof * ndof, *Kg = Gbuf + (e - e0) * ndof * ndof, *fe = fbuf + (e - e0) * ndof;
            for (int k = 0; k < nn; k++) for (int c = 0; c < 3; c++) {
                int64_t row = (int64_t)conn[e * nn + k] * 3 + c;
                if (fint) fint[row] += fe[3 * k + c];
                if (!vals) continue;
                for (int l = 0; l < nn; l++) {
                    int64_t p = find_col(colind, rowptr[row], rowptr[row + 1], 3 * conn[e * nn + l]);
                    for (int d = 0; d < 3; d++) {
                        double v = Ke[(3 * l + d) * ndof + 3 * k + c];
                        if (geometric) v += Kg[(3 * l + d) * ndof + 3 * k + c];
                        vals[p + d] += v;
                    }
                }
            }
        }
    }
    free(Kbuf); free(Gbuf); free(fbuf);
    if (vals && symmetrise) {
#pragma omp parallel for schedule(dynamic, 4096)
        for (int64_t r = 0; r < 3 * n_nodes; r++) for (int64_t p = rowptr[r]; p < rowptr[r + 1]; p++) {
            int64_t c = colind[p];
            if (c <= r) continue;
            int64_t q = find_col(colind, rowptr[c], rowptr[c + 1], (int32_t)r);
            double m = 0.5 * (vals[p] + vals[q]);
            vals[p] = m; vals[q] = m;
        }
    }
    return rc;
}

/* y = K x, CSR (the reference's matrix_vector_product is CSC K*v: src/element_assembly_structures.jl:307-309) */
void orc_spmv(int64_t n, const int64_t *rowptr, const int32_t *colind, const double *vals, const double *x, double *y) {
#pragma omp parallel for schedule(dynamic, 2048)
    for (int64_t r = 0; r < n; r++) {
        double a = 0;
        for (int64_t p = rowptr[r]; p < rowptr[r + 1]; p++) a += vals[p] * x[colind[p]];
        y[r] = a;
    }
}

/* Matrix-free internal force f_int(u) (== K u for small-strain LE), node-owner gather so that the sum
 * order per dof is fixed (elements ascending, as NodeToElementsMap src/nodal_assembly_structures.jl:69-87).
 * n2e_ptr/n2e: node -> (element*nn + local) incidences, ascending.  fixed: optional dof mask zeroed on output
 * (ext/JuliaFEMCUDAExt.jl:423-435). */
int orc_matfree(int et, int64_t n_nodes, int64_t n_elems, const double *X, const int32_t *conn, const double *u,
                int kind, const double *par, int finite_strain, const double *st_old, const uint8_t *fixed, double *y) {
    int nn = et, ndof = 3 * nn, rc = 0;
    double w[ORC_MAXGP], xi[ORC_MAXGP * 3];
    int ng = orc_quadrature(et, w, xi);
    double *fe = malloc(sizeof(double) * n_elems * ndof);
#pragma omp parallel for schedule(static) reduction(| : rc)
    for (int64_t e = 0; e < n_elems; e++) {
        double Xe[ORC_MAXN * 3], ue[ORC_MAXN * 3];
        for (int k = 0; k < nn; k++) for (int c = 0; c < 3; c++) {
            Xe[k * 3 + c] = X[(int64_t)conn[e * nn + k] * 3 + c];
            ue[k * 3 + c] = u[(int64_t)conn[e * nn + k] * 3 + c];
        }
        rc |= orc_element(et, Xe, ue, kind, par, finite_strain, 0, st_old ? st_old + e * ng * ORC_NSTATE : NULL, NULL, NULL, NULL, fe + e * ndof);
    }
    memset(y, 0, sizeof(double) * 3 * n_nodes);
    for (int64_t e = 0; e < n_elems; e++) for (int k = 0; k < nn; k++) for (int c = 0; c < 3; c++)
        y[(int64_t)conn[e * nn + k] * 3 + c] += fe[e * ndof + 3 * k + c];
    if (fixed) for (int64_t i = 0; i < 3 * n_nodes; i++) if (fixed[i]) y[i] = 0.0;
    free(fe);
    return rc;
}

/* Textbook CG exactly as src/backend/cpu.jl:221-254 / ext/JuliaFEMCUDAExt.jl:531-577 on an assembled CSR
 * operator, with the Dirichlet projection of ext:550,559 (zero fixed dofs of r and Ap).  Stop: sqrt(r.r) < tol
 * (absolute, as the reference) or, if rel != 0, sqrt(r.r) <= tol * ||b_projected||.  x is the initial guess
 * AND result (reference starts from zeros).  Returns iterations; *resid = final sqrt(r.r). */
int orc_cg_csr(int64_t n, const int64_t *rowptr, const int32_t *colind, const double *vals, const uint8_t *fixed,
               const double *b, double *x, double tol, int rel, int max_iter, double *resid) {
    double *r = malloc(sizeof(double) * n), *p = malloc(sizeof(double) * n), *Ap = malloc(sizeof(double) * n);
    orc_spmv(n, rowptr, colind, vals, x, Ap);
    double rs = 0, bn = 0;
    for (int64_t i = 0; i < n; i++) {
        int fx = fixed && fixed[i];
        r[i] = fx ? 0.0 : b[i] - Ap[i];
        bn += fx ? 0.0 : b[i] * b[i];
        p[i] = r[i]; rs += r[i] * r[i];
    }
    double thr = rel ? tol * sqrt(bn) : tol;
    int it = 0;
    if (!(sqrt(rs) < thr || (rel && sqrt(rs) <= thr))) {
        for (it = 1; it <= max_iter; it++) {
            orc_spmv(n, rowptr, colind, vals, p, Ap);
            double pAp = 0;
            for (int64_t i = 0; i < n; i++) { if (fixed && fixed[i]) Ap[i] = 0.0; pAp += p[i] * Ap[i]; }
            double a = rs / pAp, rn = 0;
            for (int64_t i = 0; i < n; i++) { x[i] += a * p[i]; r[i] -= a * Ap[i]; rn += r[i] * r[i]; }
            if (rel ? sqrt(rn) <= thr : sqrt(rn) < thr) { rs = rn; break; }
            double be = rn / rs;
            for (int64_t i = 0; i < n; i++) p[i] = r[i] + be * p[i];
            rs = rn;
        }
        if (it > max_iter) it = max_iter;
    }
    *resid = sqrt(rs);
    free(r); free(p); free(Ap);
    return it;
}

/* Greedy first-free element colouring, src/preprocess.jl:331-398 (elements visited in ascending id,
 * the deterministic stand-in for the reference's Dict iteration order). */
int orc_colouring(int et, int64_t n_nodes, int64_t n_elems, const int32_t *conn, int32_t *colour) {
    int nn = et, ncol = 0;
    int64_t *ptr = calloc(n_nodes + 1, sizeof(int64_t));
    for (int64_t i = 0; i < n_elems * nn; i++) ptr[conn[i] + 1]++;
    for (int64_t i = 0; i < n_nodes; i++) ptr[i + 1] += ptr[i];
    int64_t *fill = malloc(sizeof(int64_t) * (n_nodes + 1)), *n2e = malloc(sizeof(int64_t) * (n_elems * nn + 1));
    memcpy(fill, ptr, sizeof(int64_t) * (n_nodes + 1));
    for (int64_t e = 0; e < n_elems; e++) for (int k = 0; k < nn; k++) n2e[fill[conn[e * nn + k]]++] = e;
    for (int64_t e = 0; e < n_elems; e++) colour[e] = -1;
    uint8_t used[1024];
    for (int64_t e = 0; e < n_elems; e++) {
        memset(used, 0, sizeof used);
        for (int k = 0; k < nn; k++) {
            int32_t a = conn[e * nn + k];
            for (int64_t p = ptr[a]; p < ptr[a + 1]; p++) { int32_t c = colour[n2e[p]]; if (c >= 0 && c < 1024) used[c] = 1; }
        }
        int c = 0; while (used[c]) c++;
        colour[e] = c; if (c + 1 > ncol) ncol = c + 1;
    }
    free(ptr); free(fill); free(n2e);
    return ncol;
}

/* ------------------------------------------------------------------ external loads (f_ext) */

/* Consistent body load, src/problems_elasticity.jl:412-426: per Gauss point f_ext[3k+i] += w N_k b_i with
 * w = ip.weight * detJ ("displacement load" vector or "displacement load i" components).
 * NOTE on :418-425: as shipped, the per-component branch ASSIGNS (`f_ext[j] = b * f_buffer_dim[i]`, inner `i`
 * shadowing the component index), which would keep only the last Gauss point and always write component 1.
 * The reference's own end-to-end answer (examples/linear_static.jl:133, max|u| = 2.4052929896922337) is reproduced
 * to 6e-12 by the ACCUMULATING form below and not by the literal one (0.61 instead of 2.41), so the accumulating
 * form -- what the vector branch :412-416 does -- is the specification.  b: per-element 3 values (n_elems x 3). */
void orc_body_load(int et, int64_t n_nodes, int64_t n_elems, const double *X, const int32_t *conn, const double *b, double *f) {
    int nn = et;
    double w[ORC_MAXGP], xi[ORC_MAXGP * 3];
    int ng = orc_quadrature(et, w, xi);
    memset(f, 0, sizeof(double) * 3 * n_nodes);
    for (int64_t e = 0; e < n_elems; e++) {
        double Xe[ORC_MAXN * 3];
        for (int k = 0; k < nn; k++) for (int c = 0; c < 3; c++) Xe[k * 3 + c] = X[(int64_t)conn[e * nn + k] * 3 + c];
        for (int g = 0; g < ng; g++) {
            double N[ORC_MAXN], dN[ORC_MAXN * 3], J[9] = {0};
            orc_shape_N(et, xi + 3 * g, N);
            orc_shape_dN(et, xi + 3 * g, dN);
            for (int i = 0; i < nn; i++) for (int a = 0; a < 3; a++) for (int c = 0; c < 3; c++) J[a * 3 + c] += dN[i * 3 + a] * Xe[i * 3 + c];
            double wd = w[g] * det3(J);
            for (int k = 0; k < nn; k++) for (int c = 0; c < 3; c++) f[(int64_t)conn[e * nn + k] * 3 + c] += wd * N[k] * b[e * 3 + c];
        }
    }
}

/* Surface elements of a 3D problem: Tri3 / Tri6 / Quad4 (Elasticity3DSurfaceElements, src/problems_elasticity.jl:454-458).
 * Shape functions src/basis/lagrange_generated.jl:99-115 (Tri3), :127-143 (Tri6), :155-171 (Quad4), expanded as generated. */
static void surf_shape(int nn, double u, double v, double *N, double *dN /* nn x 2 */) {
    if (nn == 3) {
        N[0] = 1 + -1.0 * u + -1.0 * v; N[1] = u; N[2] = v;
        double d[6] = {-1, -1, 1, 0, 0, 1}; memcpy(dN, d, sizeof d);
    } else if (nn == 6) {
        N[0] = 1 + -3.0 * u + -3.0 * v + 2.0 * u * u + 4.0 * (u * v) + 2.0 * v * v;
        N[1] = -1.0 * u + 2.0 * u * u; N[2] = -1.0 * v + 2.0 * v * v;
        N[3] = 4.0 * u + -4.0 * u * u + -4.0 * (u * v); N[4] = 4.0 * (u * v); N[5] = 4.0 * v + -4.0 * (u * v) + -4.0 * v * v;
        double d[12] = {-3.0 + 2.0 * (2 * u) + 4.0 * v, -3.0 + 4.0 * u + 2.0 * (2 * v), -1.0 + 2.0 * (2 * u), 0, 0, -1.0 + 2.0 * (2 * v),
                        4.0 + -4.0 * (2 * u) + -4.0 * v, -4.0 * u, 4.0 * v, 4.0 * u, -4.0 * v, 4.0 + -4.0 * u + -4.0 * (2 * v)};
        memcpy(dN, d, sizeof d);
    } else {
        N[0] = 0.25 + -0.25 * u + -0.25 * v + 0.25 * (u * v); N[1] = 0.25 + 0.25 * u + -0.25 * v + -0.25 * (u * v);
        N[2] = 0.25 + 0.25 * u + 0.25 * v + 0.25 * (u * v); N[3] = 0.25 + -0.25 * u + 0.25 * v + -0.25 * (u * v);
        double d[8] = {-0.25 + 0.25 * v, -0.25 + 0.25 * u, 0.25 + -0.25 * v, -0.25 + -0.25 * u, 0.25 + 0.25 * v, 0.25 + 0.25 * u,
                       -0.25 + -0.25 * v, 0.25 + -0.25 * u};
        memcpy(dN, d, sizeof d);
    }
}
/* default rules (src/elements/integrate.jl:16,22-23): Tri3 GLTRI1, Tri6 GLTRI3 (src/quadrature/gltri.jl:7-27), Quad4 GLQUAD4 */
static int surf_rule(int nn, double *w, double *xi /* ng x 2 */) {
    if (nn == 3) { w[0] = 0.5; xi[0] = xi[1] = 1.0 / 3.0; return 1; }
    if (nn == 6) {
        double p[6] = {2.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0};
        memcpy(xi, p, sizeof p); w[0] = w[1] = w[2] = 1.0 / 6.0; return 3;
    }
    const double g[2] = {-0.5773502691896258, 0.5773502691896258};
    int q = 0;
    for (int j = 0; j < 2; j++) for (int i = 0; i < 2; i++) { xi[2 * q] = g[i]; xi[2 * q + 1] = g[j]; w[q] = 1.0; q++; }
    return 4;
}

/* Surface traction and pressure, src/problems_elasticity.jl:454-502: per integration point w = ip.weight * detJ with
 * detJ = || dX/dxi1 x dX/dxi2 || (src/elements/elements.jl:799-810); traction: f += w vec(T N) (:472-475);
 * pressure: n = cross(J[:,1], J[:,2]) normalised, p = -pressure, f += w p vec(n N) (:485-491).
 * faces: nn x n_faces 0-based node ids; traction: 3 per face or NULL; pressure: 1 per face or NULL.  f is ADDED to. */
void orc_surface_load(int nn, int64_t n_faces, const double *X, const int32_t *faces, const double *traction, const double *pressure, double *f) {
    double w[4], xi[8];
    int ng = surf_rule(nn, w, xi);
    for (int64_t fc = 0; fc < n_faces; fc++) {
        double fe[18] = {0};
        for (int g = 0; g < ng; g++) {
            double N[6], dN[12], t1[3] = {0}, t2[3] = {0};
            surf_shape(nn, xi[2 * g], xi[2 * g + 1], N, dN);
            for (int i = 0; i < nn; i++) for (int c = 0; c < 3; c++) {
                t1[c] += dN[2 * i] * X[(int64_t)faces[fc * nn + i] * 3 + c];
                t2[c] += dN[2 * i + 1] * X[(int64_t)faces[fc * nn + i] * 3 + c];
            }
            double n[3] = {t1[1] * t2[2] - t1[2] * t2[1], t1[2] * t2[0] - t1[0] * t2[2], t1[0] * t2[1] - t1[1] * t2[0]};
            double detJ = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            double wd = w[g] * detJ;
            for (int i = 0; i < nn; i++) for (int c = 0; c < 3; c++) {
                if (traction) fe[3 * i + c] += wd * traction[3 * fc + c] * N[i];
                if (pressure) fe[3 * i + c] += wd * (-pressure[fc]) * (n[c] / detJ) * N[i];
            }
        }
        for (int i = 0; i < nn; i++) for (int c = 0; c < 3; c++) f[(int64_t)faces[fc * nn + i] * 3 + c] += fe[3 * i + c];
    }
}

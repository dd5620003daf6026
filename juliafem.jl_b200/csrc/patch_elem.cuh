// patch_elem.cuh -- phase 1 of one element thread on the patch tables (shared by the kernels of matvec.cu and by the
// CPU-side layout check tests/hostcheck, which runs the same code on the host to validate patches.cpp without a GPU).
#pragma once
#include "common.h"
#include "elem.cuh"

namespace jf {

// field accessors of an element thread: w[k] = (offset of node k in the x tile, in doubles) | (staging entry << 16)
struct PField {
    const double *base;
    const uint32_t *w;
    JF_HD double operator()(int k, int c) const { return base[(w[k] & 0xFFFFu) + c]; }
};
struct XField4 {   // coordinates of the 4 geometry nodes of an affine element (Tet10 vertices; Hex8 nodes 0, 1, 3, 4): offsets packed two per word
    const double *base;
    const uint32_t *xw;
    JF_HD double operator()(int k, int c) const { return base[((k & 1) ? (xw[k >> 1] >> 16) : (xw[k >> 1] & 0xFFFFu)) + c]; }
};

// ---- phase 1 of one element (thread tid of a T-wide element group); et = element table of the patch
template <int NNPE, int CLS, int MODE, class Pt, int T>
JF_HD bool element_phase(const Pt &pt, long long el, const uint32_t *et, int tid, const double *xs, const double *Xs,
                                              const double *us, double *stage) {
    constexpr int NF = Pt::NF;
    uint32_t w[NNPE];
    JF_UNROLL for (int k = 0; k < NNPE; k++) w[k] = et[k * T + tid];
    auto out = [&](int k, double v0, double v1, double v2) {
        double *d = stage + 3 * (w[k] >> 16);
        d[0] = v0; d[1] = v1; d[2] = v2;
    };
    if constexpr (CLS == CLASS_AFFINE && MODE == OP_LINEAR && NNPE == 10) {
        uint32_t xw[2];
        xw[0] = et[NNPE * T + tid]; xw[1] = et[(NNPE + 1) * T + tid];
        PField U{xs, w};
        XField4 X{Xs, xw};
        Pt q = pt;
        q.load(el);
        tet10_affine_linear(q.la, q.mu, U, X, out);
        return true;
    } else if constexpr (CLS == CLASS_AFFINE && MODE == OP_LINEAR && NNPE == 8) {
        uint32_t xw[2];
        xw[0] = et[NNPE * T + tid]; xw[1] = et[(NNPE + 1) * T + tid];
        PField U{xs, w};
        XField4 X{Xs, xw};
        Pt q = pt;
        q.load(el);
        hex8_affine_linear(q.la, q.mu, U, X, out);
        return true;
    } else {
        PField X{Xs, w};
        PField F[NF];
        F[0].base = xs; F[0].w = w;
        if (NF == 2) { F[NF - 1].base = us; F[NF - 1].w = w; }
        if (NNPE == 10) return tet10_general(pt, el, F, X, out);
        else if (NNPE == 8) return hex8_general(pt, el, F, X, out);
        else return tet4_general(pt, el, F, X, out);
    }
}

}  // namespace jf

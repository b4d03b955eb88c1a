// zkm_curve.cuh -- short-Weierstrass (a = 0) group law in XYZZ coordinates,
// generic over the coordinate field (Fq for G1, Fq2 for G2).
//
// Replaces the Jacobian arithmetic of un-vendored ark-ec 0.3.0
// src/models/short_weierstrass_jacobian.rs (GroupProjective::add_assign_mixed,
// add_assign, double_in_place, into_affine).  A different
// projective system is legal because parity is defined on the normalised affine
// point, which is unique.  Formulas: EFD madd-2008-s, add-2008-s, dbl-2008-s-1,
// mdbl-2008-s-1; identity is encoded as ZZ = 0.  All exceptional cases (identity
// operands, P + P, P + (-P)) are handled exactly -- proving keys do contain
// repeated points and points at infinity (SURVEY.md section 7 "hard parts").
#pragma once
#include "zkm_field.cuh"

namespace zkm {

template <class F>
struct Affine {
    F x, y;
};

template <class F>
struct XYZZ {
    F X, Y, ZZ, ZZZ;
    static ZKM_DEV XYZZ identity() {
        XYZZ r;
        r.X = F::zero();
        r.Y = F::zero();
        r.ZZ = F::zero();
        r.ZZZ = F::zero();
        return r;
    }
    ZKM_DEV bool is_identity() const { return ZZ.is_zero(); }
};

template <class F>
ZKM_DEV XYZZ<F> xyzz_from_affine(const F& x, const F& y) {
    XYZZ<F> r;
    r.X = x;
    r.Y = y;
    r.ZZ = F::one();
    r.ZZZ = F::one();
    return r;
}

// 2 * (x, y) for an affine point (mdbl-2008-s-1).  y = 0 gives the identity.
template <class F>
#if !defined(ZKM_HOST_EMU)
__device__ __noinline__
#endif
void xyzz_mdbl(XYZZ<F>& r, const F& x, const F& y) {
    F U = dbl(y);
    F V = sqr(U);
    F W = U * V;
    F S = x * V;
    F xx = sqr(x);
    F M = dbl(xx) + xx;
    F X3 = sqr(M) - dbl(S);
    r.Y = M * (S - X3) - W * y;
    r.X = X3;
    r.ZZ = V;
    r.ZZZ = W;
}

// p = 2 * p  (dbl-2008-s-1)
template <class F>
ZKM_DEV void xyzz_dbl(XYZZ<F>& p) {
    if (p.is_identity()) return;
    F U = dbl(p.Y);
    F V = sqr(U);
    F W = U * V;
    F S = p.X * V;
    F xx = sqr(p.X);
    F M = dbl(xx) + xx;
    F X3 = sqr(M) - dbl(S);
    p.Y = M * (S - X3) - W * p.Y;
    p.X = X3;
    p.ZZ = V * p.ZZ;
    p.ZZZ = W * p.ZZZ;
}

// p += (x2, y2), the affine operand is NOT the identity  (madd-2008-s, 8M + 2S)
template <class F>
ZKM_DEV void xyzz_madd(XYZZ<F>& p, const F& x2, const F& y2) {
    if (p.is_identity()) {
        p = xyzz_from_affine(x2, y2);
        return;
    }
    F Pv = x2 * p.ZZ - p.X;
    F R = y2 * p.ZZZ - p.Y;
    if (Pv.is_zero()) {
        if (R.is_zero())
            xyzz_mdbl(p, x2, y2);
        else
            p = XYZZ<F>::identity();
        return;
    }
    F PP = sqr(Pv);
    F PPP = Pv * PP;
    F Q = p.X * PP;
    F X3 = (sqr(R) - PPP) - dbl(Q);
    p.Y = R * (Q - X3) - p.Y * PPP;
    p.X = X3;
    p.ZZ = p.ZZ * PP;
    p.ZZZ = p.ZZZ * PPP;
}

// p += q  (add-2008-s, 12M + 2S)
template <class F>
ZKM_DEV void xyzz_add(XYZZ<F>& p, const XYZZ<F>& q) {
    if (q.is_identity()) return;
    if (p.is_identity()) {
        p = q;
        return;
    }
    F U1 = p.X * q.ZZ;
    F S1 = p.Y * q.ZZZ;
    F Pv = q.X * p.ZZ - U1;
    F R = q.Y * p.ZZZ - S1;
    if (Pv.is_zero()) {
        if (R.is_zero())
            xyzz_dbl(p);
        else
            p = XYZZ<F>::identity();
        return;
    }
    F PP = sqr(Pv);
    F PPP = Pv * PP;
    F Q = U1 * PP;
    F X3 = (sqr(R) - PPP) - dbl(Q);
    p.Y = R * (Q - X3) - S1 * PPP;
    p.X = X3;
    p.ZZ = (p.ZZ * q.ZZ) * PP;
    p.ZZZ = (p.ZZZ * q.ZZZ) * PPP;
}

// into_affine(): x = X/ZZ, y = Y/ZZZ.  Returns false for the identity.
template <class F>
ZKM_DEV bool xyzz_to_affine(const XYZZ<F>& p, F& x, F& y) {
    if (p.is_identity()) return false;
    F i = inv(p.ZZ * p.ZZZ);
    x = (p.X * p.ZZZ) * i;
    y = (p.Y * p.ZZ) * i;
    return true;
}

// k * (x, y) by left-to-right double-and-add (input generation / tests only).
template <class F>
ZKM_DEV XYZZ<F> xyzz_mul_u64(const F& x, const F& y, uint64_t k) {
    XYZZ<F> r = XYZZ<F>::identity();
    for (int b = 63; b >= 0; b--) {
        xyzz_dbl(r);
        if ((k >> b) & 1) xyzz_madd(r, x, y);
    }
    return r;
}

}  // namespace zkm

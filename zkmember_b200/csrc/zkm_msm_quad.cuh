// zkm_msm_quad.cuh -- lane-cooperative XYZZ group law for the serial tails of the MSM.
//
// The window reduction, the fold levels of a small MSM and the Horner combine are chains of dependent
// group operations on a handful of threads: pure latency (a dependent Montgomery product costs ~1.8 k
// cycles on a lone warp).  Here FOUR lanes hold identical copies of the operands; the products of a formula
// are scheduled by dependency depth (dbl-2008-s-1: 9 products, depth 3; add-2008-s: 14 products, depth 4),
// every level is ONE product per lane on lane-selected operands (no divergence) and shuffles hand the
// results back to all four lanes.  Same values as xyzz_dbl / xyzz_add, including the exceptional cases.
#pragma once
#include "zkm_msm.cuh"

namespace zkm {

template <class P>
__device__ __forceinline__ Fp<P> quad_bcast(const Fp<P>& v, int src, uint32_t mask) {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < P::N; i++) r.l[i] = __shfl_sync(mask, v.l[i], src, 4);
    return r;
}
template <class P>
__device__ __forceinline__ Fp2<P> quad_bcast(const Fp2<P>& v, int src, uint32_t mask) {
    Fp2<P> r;
    r.c0 = quad_bcast(v.c0, src, mask);
    r.c1 = quad_bcast(v.c1, src, mask);
    return r;
}
template <class P>
__device__ __forceinline__ Fp<P> quad_sel(int q, const Fp<P>& a0, const Fp<P>& a1, const Fp<P>& a2, const Fp<P>& a3) {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < P::N; i++) r.l[i] = q == 0 ? a0.l[i] : (q == 1 ? a1.l[i] : (q == 2 ? a2.l[i] : a3.l[i]));
    return r;
}
template <class P>
__device__ __forceinline__ Fp2<P> quad_sel(int q, const Fp2<P>& a0, const Fp2<P>& a1, const Fp2<P>& a2, const Fp2<P>& a3) {
    Fp2<P> r;
    r.c0 = quad_sel(q, a0.c0, a1.c0, a2.c0, a3.c0);
    r.c1 = quad_sel(q, a0.c1, a1.c1, a2.c1, a3.c1);
    return r;
}

// p = 2 p; p is identical on the four lanes q = 0..3 of `mask` on entry and on exit
template <class F>
__device__ __forceinline__ void xyzz_dbl_quad_inl(XYZZ<F>& p, int q, uint32_t mask) {
    if (p.is_identity()) return;
    const F U = dbl(p.Y);
    F r = quad_sel(q, U, p.X, U, U);
    r = r * r;                                    // lane 0: V = U^2, lane 1: XX = X^2
    const F V = quad_bcast(r, 0, mask), XX = quad_bcast(r, 1, mask);
    const F M = dbl(XX) + XX;
    r = quad_sel(q, U, p.X, p.ZZ, M) * quad_sel(q, V, V, V, M);   // W = U V | S = X V | ZZ' = ZZ V | M^2
    const F Wv = quad_bcast(r, 0, mask), S = quad_bcast(r, 1, mask), ZZ3 = quad_bcast(r, 2, mask), MM = quad_bcast(r, 3, mask);
    const F X3 = MM - dbl(S);
    r = quad_sel(q, M, Wv, Wv, Wv) * quad_sel(q, S - X3, p.Y, p.ZZZ, p.ZZZ);   // M (S - X3) | W Y | ZZZ' = W ZZZ
    const F T1 = quad_bcast(r, 0, mask), T2 = quad_bcast(r, 1, mask), ZZZ3 = quad_bcast(r, 2, mask);
    p.X = X3;
    p.Y = T1 - T2;
    p.ZZ = ZZ3;
    p.ZZZ = ZZZ3;
}

// out-of-line copy for the kernels that call it from several sites (code size, ptxas time)
template <class F>
__device__ __noinline__ void xyzz_dbl_quad(XYZZ<F>& p, int q, uint32_t mask) { xyzz_dbl_quad_inl(p, q, mask); }

// p += s; both identical on the four lanes
template <class F>
__device__ __noinline__ void xyzz_add_quad(XYZZ<F>& p, const XYZZ<F>& s, int q, uint32_t mask) {
    if (s.is_identity()) return;
    if (p.is_identity()) {
        p = s;
        return;
    }
    F r = quad_sel(q, p.X, s.X, p.Y, s.Y) * quad_sel(q, s.ZZ, p.ZZ, s.ZZZ, p.ZZZ);   // U1 | U2 | S1 | S2
    const F U1 = quad_bcast(r, 0, mask), U2 = quad_bcast(r, 1, mask), S1 = quad_bcast(r, 2, mask), S2 = quad_bcast(r, 3, mask);
    const F Pv = U2 - U1, R = S2 - S1;
    if (Pv.is_zero()) {
        if (R.is_zero()) xyzz_dbl_quad(p, q, mask);
        else p = XYZZ<F>::identity();
        return;
    }
    r = quad_sel(q, Pv, R, p.ZZ, p.ZZZ) * quad_sel(q, Pv, R, s.ZZ, s.ZZZ);           // PP | RR | ZZ1 ZZ2 | ZZZ1 ZZZ2
    const F PP = quad_bcast(r, 0, mask), RR = quad_bcast(r, 1, mask), ZZ12 = quad_bcast(r, 2, mask), ZZZ12 = quad_bcast(r, 3, mask);
    r = quad_sel(q, Pv, U1, ZZ12, ZZ12) * PP;                                        // PPP | Q | ZZ3
    const F PPP = quad_bcast(r, 0, mask), Q = quad_bcast(r, 1, mask), ZZ3 = quad_bcast(r, 2, mask);
    const F X3 = (RR - PPP) - dbl(Q);
    r = quad_sel(q, R, S1, ZZZ12, ZZZ12) * quad_sel(q, Q - X3, PPP, PPP, PPP);       // R (Q - X3) | S1 PPP | ZZZ3
    const F T1 = quad_bcast(r, 0, mask), T2 = quad_bcast(r, 1, mask), ZZZ3 = quad_bcast(r, 2, mask);
    p.X = X3;
    p.Y = T1 - T2;
    p.ZZ = ZZ3;
    p.ZZZ = ZZZ3;
}

}  // namespace zkm

// zkm_msm_affine.cuh -- batched-affine pairwise reduction of the bucket lists (K4, large MSMs).
//
// Instead of chaining XYZZ mixed additions (10 field products per point), the sorted bucket lists are
// halved level by level with AFFINE additions: out[j] = in[2j] + in[2j+1] inside every bucket.  All
// additions of a level are independent, so their denominators (x1 - x0, or 2 y0 for a doubling) are
// inverted together with Montgomery's trick, two storeys high:
//
//   k_pair_fwd   thread t walks m consecutive outputs: d_e, exclusive prefix products pre[e] -> HBM,
//                thread total T[t]
//   k_inv_batch  thread u owns m2 consecutive totals: prefix products, ONE Fermat inversion, unwind ->
//                T[t] := 1 / T[t]
//   k_pair_bwd   thread t unwinds its range backwards: 1/d_e = run * pre[e], run *= d_e, then
//                lambda = (y1 - y0) / d, x3 = lambda^2 - x0 - x1, y3 = lambda (x0 - x3) - y0 -> next level
//
// = 6 field products per addition + 1/(m m2) of an inversion, against 10 for the XYZZ chain; the price is
// HBM traffic (~0.5 KB per addition, far below the roofline of this integer-bound kernel).  Exceptional
// pairs are exact: P + P doubles (denominator 2y), P + (-P) yields the identity marker, identity
// operands pass the other point through.  Identity marker in the affine arrays: top limb of x all ones
// (no reduced field element has it).  After a few levels the short remaining lists go to the XYZZ
// task kernels.  The affine result of each bucket is unique, so the final MSM bytes do not change.
#pragma once
#include "zkm_msm.cuh"

namespace zkm {

template <class P> __device__ __forceinline__ bool aff_is_identity(const Fp<P>& x) { return x.l[P::N - 1] == 0xffffffffu; }
template <class P> __device__ __forceinline__ bool aff_is_identity(const Fp2<P>& x) { return x.c0.l[P::N - 1] == 0xffffffffu; }
template <class P> __device__ __forceinline__ void aff_set_identity(Fp<P>& x, Fp<P>& y) {
#pragma unroll
    for (int i = 0; i < P::N; i++) { x.l[i] = 0xffffffffu; y.l[i] = 0; }
}
template <class P> __device__ __forceinline__ void aff_set_identity(Fp2<P>& x, Fp2<P>& y) {
    aff_set_identity(x.c0, y.c0);
    x.c1 = Fp<P>::zero();
    y.c1 = Fp<P>::zero();
}

// position of input i of the current level: level 0 reads the sorted (index | sign) list and gathers the
// registered bases; later levels read the affine array written by the previous level
template <class F, bool L0>
__device__ __forceinline__ const char* pair_src(const char* src, const uint32_t* idx, uint32_t i, uint32_t& sign) {
    constexpr int CB = CoordIO<F>::BYTES;
    if (L0) {
        uint32_t id = idx[i];
        sign = id >> 31;
        return src + (size_t)(id & 0x7fffffffu) * (2 * CB);
    }
    sign = 0;
    return src + (size_t)i * (2 * CB);
}

// Work distribution of a level (round 2): the E outputs are cut into blocks of 32 m consecutive outputs, one block per
// WARP; lane l of the warp owns the chain of outputs l, l + 32, l + 64, ... of the block (m of them).  Consecutive
// lanes therefore touch consecutive outputs at every step: the map words, the prefix products, the outputs and -- from
// level 1 on -- the inputs are read and written as contiguous 32-element runs.  (Round 1 gave every thread m
// CONSECUTIVE outputs: each warp-wide access touched 32 different lines 6 KB apart; the forward pass of levels >= 1 ran
// at 2.2 TB/s and the unwind pass lost 27 % of its stall cycles to memory.)  Montgomery's trick does not care which
// elements share a chain.  Where output e finds its inputs is precomputed per level by k_pair_map (zkm_msm.cu):
// map[e] = first input position | (two inputs ? 1 << 31 : 0).
__device__ __forceinline__ uint32_t pair_chains(uint32_t E, uint32_t m) { return 32u * ((E + 32u * m - 1u) / (32u * m)); }

// classification of one output element
struct PairKind {
    bool pair;      // two inputs (else: the odd element of the list is passed through)
    bool has_d;     // contributes a denominator to the batch
    bool dbl;       // P + P
    bool cancel;    // P + (-P)
    bool inf0, inf1;
};

// Level 0 of the forward pass only needs the x coordinates.  A 48-byte x inside a 96-byte record that is merely
// 32-byte aligned straddles two 64-byte DRAM bursts half of the time (ncu: 29 GB read for 14 GB of x); gathered
// from a separate array of 64-byte-aligned slots every x costs exactly one burst.  XSLOT = 0: no such array.
template <class F> struct XArr { static constexpr int SLOT = (CoordIO<F>::BYTES == 48) ? 64 : 0; };

template <class F>
__global__ void __launch_bounds__(256) k_build_xarr(const char* __restrict__ bases, uint64_t n, char* __restrict__ xarr) {
    constexpr int CB = CoordIO<F>::BYTES;
    constexpr int SLOT = XArr<F>::SLOT > 0 ? XArr<F>::SLOT : CB;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        CoordIO<F>::st(xarr + i * SLOT, CoordIO<F>::ld(bases + i * 2 * CB));
}

template <class F, bool L0>
__global__ void __launch_bounds__(256)
k_pair_fwd(const char* __restrict__ src, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ map,
           const uint32_t* __restrict__ off_out, uint32_t K, uint32_t m, char* __restrict__ pre, char* __restrict__ T,
           const char* __restrict__ xarr) {
    constexpr int CB = CoordIO<F>::BYTES;
    constexpr int XS = XArr<F>::SLOT > 0 ? XArr<F>::SLOT : 2 * CB;
    const uint32_t E = off_out[K];
    const bool use_x = L0 && XArr<F>::SLOT > 0 && xarr != nullptr;
    const uint32_t nT = pair_chains(E, m);
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nT; t += gridDim.x * blockDim.x) {
        const uint32_t e0 = (t >> 5) * 32u * m + lane;
        F run = F::one();
        for (uint32_t j = 0; j < m; j++) {
            const uint32_t e = e0 + 32u * j;
            if (e >= E) break;
            const uint32_t mp = map[e];
            CoordIO<F>::st(pre + (size_t)e * CB, run);
            if (mp >> 31) {
                const uint32_t i0 = mp & 0x7fffffffu;
                uint32_t s0, s1;
                const char* p0 = pair_src<F, L0>(src, idx, i0, s0);
                const char* p1 = pair_src<F, L0>(src, idx, i0 + 1, s1);
                F x0, x1;
                if (use_x) {
                    x0 = CoordIO<F>::ld_gather(xarr + (size_t)(idx[i0] & 0x7fffffffu) * XS);
                    x1 = CoordIO<F>::ld_gather(xarr + (size_t)(idx[i0 + 1] & 0x7fffffffu) * XS);
                } else {
                    x0 = L0 ? CoordIO<F>::ld_gather(p0) : CoordIO<F>::ld_plain(p0);
                    x1 = L0 ? CoordIO<F>::ld_gather(p1) : CoordIO<F>::ld_plain(p1);
                }
                if (!aff_is_identity(x0) && !aff_is_identity(x1)) {
                    if (x0 != x1) {
                        run = run * (x1 - x0);
                    } else {
                        F y0 = L0 ? CoordIO<F>::ld_gather(p0 + CB) : CoordIO<F>::ld_plain(p0 + CB);
                        F y1 = L0 ? CoordIO<F>::ld_gather(p1 + CB) : CoordIO<F>::ld_plain(p1 + CB);
                        if (s0) y0 = neg(y0);
                        if (s1) y1 = neg(y1);
                        if (y0 == y1) run = run * dbl(y0);
                    }
                }
            }
        }
        CoordIO<F>::st(T + (size_t)t * CB, run);
    }
}

// T[i] := 1 / T[i] for i < nT = pair_chains(E, m)
template <class F>
__global__ void __launch_bounds__(128)
k_inv_batch(const uint32_t* __restrict__ off_out, uint32_t K, uint32_t m, uint32_t m2, char* __restrict__ T,
            char* __restrict__ pre2) {
    constexpr int CB = CoordIO<F>::BYTES;
    const uint32_t E = off_out[K];
    const uint32_t nT = pair_chains(E, m);
    const uint32_t nU = (nT + m2 - 1) / m2;
    for (uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < nU; u += gridDim.x * blockDim.x) {
        const uint32_t i0 = u * m2, i1 = (i0 + m2 < nT) ? i0 + m2 : nT;
        F run = F::one();
        for (uint32_t i = i0; i < i1; i++) {
            CoordIO<F>::st(pre2 + (size_t)i * CB, run);
            run = run * CoordIO<F>::ld_plain(T + (size_t)i * CB);
        }
        F r = inv(run);
        for (uint32_t i = i1; i-- > i0;) {
            F t = CoordIO<F>::ld_plain(T + (size_t)i * CB);
            CoordIO<F>::st(T + (size_t)i * CB, r * CoordIO<F>::ld_plain(pre2 + (size_t)i * CB));
            r = r * t;
        }
    }
}

// CTAs per SM the unwind kernel is compiled for: 4 (<= 128 registers) for the 256/384-bit fields; the 761-bit
// field of BW6-761 keeps ~6 coordinates of 24 limbs live and gets 2 (<= 255 registers) instead of spilling.
template <class F> struct PairBwdMinBlocks { static constexpr int value = 4; };
template <> struct PairBwdMinBlocks<Bw6_761_Fq> { static constexpr int value = 2; };
// round 2 (after the strided chains freed the cursor registers), measured at 2^24: 6 CTAs/SM for the 254-bit field (80 registers,
// 32 bytes of stack) -2 % on the levels; 5 for the 381-bit field (96 registers, 128 bytes of stack) +3 % -- it stays at 4.
template <> struct PairBwdMinBlocks<Bn254_Fq> { static constexpr int value = 6; };

template <class F, bool L0>
__global__ void __launch_bounds__(128, PairBwdMinBlocks<F>::value)
k_pair_bwd(const char* __restrict__ src, const uint32_t* __restrict__ idx, const uint32_t* __restrict__ map,
           const uint32_t* __restrict__ off_out, uint32_t K, uint32_t m, const char* __restrict__ pre,
           const char* __restrict__ Tinv, char* __restrict__ dst) {
    constexpr int CB = CoordIO<F>::BYTES;
    const uint32_t E = off_out[K];
    const uint32_t nT = pair_chains(E, m);
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < nT; t += gridDim.x * blockDim.x) {
        const uint32_t e0 = (t >> 5) * 32u * m + lane;
        if (e0 >= E) continue;
        uint32_t cnt = (E - e0 + 31u) >> 5;     // outputs of this chain
        if (cnt > m) cnt = m;
        F run = CoordIO<F>::ld_plain(Tinv + (size_t)t * CB);
        for (uint32_t j = cnt; j-- > 0;) {
            const uint32_t e = e0 + 32u * j;
            const uint32_t mp = map[e];
            const uint32_t i0 = mp & 0x7fffffffu;
            uint32_t s0, s1;
            const char* p0 = pair_src<F, L0>(src, idx, i0, s0);
            F x0 = L0 ? CoordIO<F>::ld_gather(p0) : CoordIO<F>::ld_plain(p0);
            F y0 = L0 ? CoordIO<F>::ld_gather(p0 + CB) : CoordIO<F>::ld_plain(p0 + CB);
            if (s0) y0 = neg(y0);
            F x3 = x0, y3 = y0;
            if (mp >> 31) {
                const char* p1 = pair_src<F, L0>(src, idx, i0 + 1, s1);
                F x1 = L0 ? CoordIO<F>::ld_gather(p1) : CoordIO<F>::ld_plain(p1);
                F y1 = L0 ? CoordIO<F>::ld_gather(p1 + CB) : CoordIO<F>::ld_plain(p1 + CB);
                if (s1) y1 = neg(y1);
                const bool inf0 = aff_is_identity(x0), inf1 = aff_is_identity(x1);
                if (inf0) {
                    x3 = x1;
                    y3 = y1;
                } else if (!inf1) {
                    const bool same_x = (x0 == x1);
                    if (same_x && y0 != y1) {
                        aff_set_identity(x3, y3);   // P + (-P)
                    } else {
                        F d = same_x ? dbl(y0) : (x1 - x0);
                        F dinv = run * CoordIO<F>::ld_plain(pre + (size_t)e * CB);
                        run = run * d;
                        F num;
                        if (same_x) {
                            F xx = sqr(x0);
                            num = dbl(xx) + xx;
                        } else {
                            num = y1 - y0;
                        }
                        F lam = num * dinv;
                        x3 = sqr(lam) - x0 - x1;
                        y3 = lam * (x0 - x3) - y0;
                    }
                }
            }
            CoordIO<F>::st(dst + (size_t)e * (2 * CB), x3);
            CoordIO<F>::st(dst + (size_t)e * (2 * CB) + CB, y3);
        }
    }
}

}  // namespace zkm

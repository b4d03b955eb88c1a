// zkm_fpmul_u.cuh -- Montgomery product on an UNSATURATED radix: every partial product is one plain
// `mad.wide.u32` (IMAD.WIDE.U32 with a 64-bit addend and NO carry predicate).
//
// Why: on sm_100a the carry-chained forms the saturated 32-bit CIOS needs (IMAD.WIDE.U32.X with a predicate
// carry in/out, IMAD.HI) issue at HALF the rate of a plain IMAD.WIDE (tools/microbench/imad_peak.cu on B200:
// 62 vs 31 per clock per SM, profiles/imad_peak_r1.jsonl).  With limbs of r <= 30 bits a column
// sum_{i+j=k} a_i b_j of up to 15 products (each < 2^60) fits a 64-bit accumulator, so the products need no
// carries at all; carries are resolved once per COLUMN with shifts and adds on the otherwise idle ALU pipe.
//
// Interface and value are unchanged: inputs and output are saturated 32-bit limbs in Montgomery form with
// R = 2^(32 N), the result is a * b * R^-1 mod p fully reduced -- byte-identical to the CIOS product
// (fp_mul_cios, kept as the cross-check) and therefore to ark-ff 0.3.0's Fp::mul_assign.
//
// Schedule (product scanning, columns of r bits; UM = ceil(BITS / r) operand limbs, MM = ceil(32 N / r)
// reduction limbs, F = floor(32 N / r), t = 32 N - r F):
//   column k:  A = carry + sum_{i+j=k} a_i b_j                     (<= UM products, < 2^64)
//              B = (A mod 2^r) + sum_{j<k} m_j p_{k-j}             (<= MM products)
//              k < MM:  m_k = (B * (-p^-1)) mod 2^r  (mod 2^t in the last reduction column when t > 0),
//                       B += m_k p_0                               (low r -- or t -- bits of B are now zero)
//              carry = (A >> r) + (B >> r);   k >= F: limb_{k-F} = B mod 2^r
// The reduction removes exactly 32 N bits (F full columns and t bits of column F), so the quotient is
// (a b + m p) / 2^(32 N) < 2 p with m < 2^(32 N): the same value the 32-bit CIOS produces before its final
// subtraction.  The r-bit result limbs are re-packed into 32-bit words (offset t) and reduced once.
#pragma once
#include "zkm_arith.cuh"

namespace zkm {

template <class P>
struct FpU {   // an operand split into UM limbs of UR bits
    static constexpr int UR = P::UR;
    static constexpr int UM = (P::BITS + UR - 1) / UR;
    uint32_t u[UM];
};

#if defined(ZKM_HOST_EMU)
inline uint32_t zkm_funnel_r(uint32_t lo, uint32_t hi, int s) { return s ? (uint32_t)((((uint64_t)hi << 32) | lo) >> s) : lo; }
#else
ZKM_DEV uint32_t zkm_funnel_r(uint32_t lo, uint32_t hi, int s) { return __funnelshift_r(lo, hi, s); }
#endif

// limbs of r bits out of N saturated 32-bit words (a < 2^BITS)
template <class P>
ZKM_DEV FpU<P> fp_split(const uint32_t (&l)[P::N]) {
    constexpr int N = P::N, r = P::UR, UM = FpU<P>::UM;
    constexpr uint32_t mask = (1u << r) - 1u;
    FpU<P> o;
    ZKM_UNROLL
    for (int i = 0; i < UM; i++) {
        const int b0 = r * i, w0 = b0 >> 5, s = b0 & 31;
        const uint32_t lo = l[w0];
        const uint32_t hi = (w0 + 1 < N) ? l[w0 + 1] : 0u;
        o.u[i] = zkm_funnel_r(lo, hi, s) & mask;
    }
    return o;
}

template <class P>
struct FpUCfg {
    static constexpr int N = P::N, r = P::UR;
    static constexpr int UM = FpU<P>::UM;
    static constexpr int MM = (32 * N + r - 1) / r;
    static constexpr int F = (32 * N) / r;
    static constexpr int T = 32 * N - r * F;
    static constexpr int KA = 2 * UM - 1, KB = MM + UM - 1;
    static constexpr int KC = (KA > KB ? KA : KB) + 1;     // + the column that receives the last carry
    static constexpr int NL = KC - F;                      // result limbs
    static_assert(r <= 30 && r >= 24, "radix");
    static_assert((uint64_t)(UM > MM ? UM : MM) <= ((~0ull - (1ull << 36)) >> (2 * r)), "column sums must fit 64 bits");
    static_assert(r * NL >= 32 * N + T, "result limbs cover the quotient");
};

// Products are written as `acc += (uint64_t)x * y` (NVVM emits mul.wide.u32 + add.s64, ptxas fuses them into ONE
// IMAD.WIDE.U32 chained on the accumulator pair).  Two things defeat that and are avoided here (checked in SASS):
//  * m_k is derived from a masked 64-bit value, so NVVM would turn m_k * constant into a 64-bit mul.lo.s64 and
//    ptxas would add a high-word fix-up IADD3 per product -> m_k goes through an opaque register move;
//  * explicit mad.wide.u32 inline PTX makes ptxas split every product from its addition again
//    (IMAD.WIDE ..., RZ + IADD3/IADD3.X pairs: twice the instructions).
#if defined(ZKM_HOST_EMU)
inline uint32_t zkm_opaque(uint32_t x) { return x; }
#else
ZKM_DEV uint32_t zkm_opaque(uint32_t x) {
    asm("" : "+r"(x));
    return x;
}
#endif

// the reduction half + re-packing: `colA(k, acc)` returns acc + sum_{i+j=k} a_i b_j.
// Carries between columns cost two shifts: the high WORD of the a b column sum A re-enters the next column
// through one more IMAD.WIDE (x 2^(32-r)), the reduction chain B continues on A's low word, and B >> r is
// the next column's initial accumulator.
template <class P, class ColA>
ZKM_DEV void fp_mont_columns(uint32_t (&out)[P::N], ColA colA) {
    typedef FpUCfg<P> C;
    constexpr int N = C::N, r = C::r, UM = C::UM, MM = C::MM, F = C::F, T = C::T, KC = C::KC, NL = C::NL;
    constexpr uint32_t mask = (1u << r) - 1u;
    uint32_t m[MM];
    uint32_t lim[NL + 2];
    uint64_t carry = 0;     // B >> r of the previous column
    uint32_t ahi = 0;       // high word of the previous column's A
    ZKM_UNROLL
    for (int k = 0; k < KC; k++) {
        uint64_t A = carry;
        if (k > 0) A += (uint64_t)ahi << (32 - r);
        if (k < C::KA) A = colA(k, A);
        ahi = (uint32_t)(A >> 32);
        uint64_t B = (uint32_t)A;
        ZKM_UNROLL
        for (int j = 0; j < MM; j++) {
            const int i = k - j;
            if (j < k && i >= 1 && i < UM) B += (uint64_t)m[j] * P::modu(i);
        }
        if (k < MM) {
            constexpr uint32_t tmask = (T == 0) ? mask : ((1u << (T ? T : 1)) - 1u);
            const uint32_t mk = zkm_opaque(((uint32_t)B * P::INV) & ((k == MM - 1) ? tmask : mask));
            m[k] = mk;
            B += (uint64_t)mk * P::modu(0);
        }
        carry = B >> r;
        if (k >= F) lim[k - F] = (uint32_t)B & mask;
    }
    lim[NL] = 0;
    lim[NL + 1] = 0;
    // word w = bits [32 w + T, 32 w + T + 32) of sum lim[j] 2^(r j)
    ZKM_UNROLL
    for (int w = 0; w < N; w++) {
        const int b0 = 32 * w + T, j0 = b0 / r, s = b0 % r;
        uint32_t v = lim[j0] >> s;
        v |= lim[j0 + 1] << (r - s);
        if (2 * r - s < 32) v |= lim[j0 + 2] << (2 * r - s);
        out[w] = v;
    }
}

template <class P>
ZKM_DEV void fp_mul_u_raw(uint32_t (&out)[P::N], const FpU<P>& a, const FpU<P>& b) {
    constexpr int UM = FpU<P>::UM;
    fp_mont_columns<P>(out, [&](int k, uint64_t A) {
        ZKM_UNROLL
        for (int i = 0; i < UM; i++) {
            const int j = k - i;
            if (j >= 0 && j < UM) A += (uint64_t)a.u[i] * b.u[j];
        }
        return A;
    });
}

// squaring: cross products once, against the doubled operand (2 a_i < 2^(r+1): a column of <= UM/2 doubled
// products + one square stays below 2^64 for UM <= 14 at r = 30)
template <class P>
ZKM_DEV void fp_sqr_u_raw(uint32_t (&out)[P::N], const FpU<P>& a) {
    constexpr int UM = FpU<P>::UM;
    static_assert((uint64_t)(UM / 2) * 2 + 1 <= ((~0ull - (1ull << 36)) >> (2 * P::UR)), "squaring column fits 64 bits");
    uint32_t a2[UM];
    ZKM_UNROLL
    for (int i = 0; i < UM; i++) a2[i] = a.u[i] << 1;
    fp_mont_columns<P>(out, [&](int k, uint64_t A) {
        ZKM_UNROLL
        for (int i = 0; i < UM; i++) {
            const int j = k - i;
            if (j > i && j < UM) A += (uint64_t)a2[i] * a.u[j];
        }
        if ((k & 1) == 0 && (k >> 1) < UM) A += (uint64_t)a.u[k >> 1] * a.u[k >> 1];
        return A;
    });
}

}  // namespace zkm

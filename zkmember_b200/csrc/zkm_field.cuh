// zkm_field.cuh -- Montgomery prime fields on 32-bit limbs and the quadratic
// extension Fq2 = Fq[u]/(u^2+1) used by G2 of both BLS12-381 and BN254 (BW6-761's G2 is over Fq itself).
//
// Semantics replaced (un-vendored ark-ff 0.3.0, pinned at
// /root/reference/Cargo.lock:229-230): Fp256/Fp384 mul_assign / square_in_place /
// add_assign / sub_assign / neg (src/fields/macros.rs, arithmetic.rs) and
// QuadExtField (src/fields/models/quadratic_extension.rs).  Values are value*R mod p
// with R = 2^(32*N) = 2^(64*limbs64): the byte image of an element is identical to
// ark-ff's little-endian u64-limb storage, and every operation returns the fully
// reduced representative, so results are bit-identical to ark-ff's.
//
// The Montgomery product is an operand-scanning CIOS on ABSOLUTE columns with two
// accumulators: acc[0] takes every carry chain that starts on an even column,
// acc[1] every chain that starts on an odd column.  A chain is lo(a_i*b_j) into
// column k, hi(a_i*b_j) into column k+1, for i stepping by two -- i.e. exactly the
// mad.lo.cc / madc.hi.cc pairing that ptxas turns into one IMAD.WIDE.U32.X on an
// aligned register pair.  Column j of the two accumulators is merged right before
// the reduction factor m_j is formed; the carry of a chain lands in a column no
// chain body has reached yet, so it can never overflow.  Requires a spare top bit
// in the modulus (381/384, 255/256, 254/256, 761/768, 377/384 here).
#pragma once
#include "zkm_arith.cuh"
#include "zkm_constants.cuh"
#include "zkm_fpmul_u.cuh"

namespace zkm {

// ----------------------------------------------------------------------------- parameter tags
#define ZKM_DEFINE_FP_PARAMS(Tag, PREFIX, NLIMBS)                                   \
    struct Tag {                                                                    \
        static constexpr int N = NLIMBS;                                            \
        static constexpr uint32_t INV = PREFIX##_INV32;                             \
        static ZKM_DEV uint32_t inv_rt() { return PREFIX##_INVC[0]; }  /* INV as a run-time value (see gen_constants.py) */ \
        static constexpr int BITS = PREFIX##_BITS;                                  \
        static constexpr int UR = PREFIX##_UR;   /* radix bits of the carry-free product */ \
        static ZKM_CEXPR uint32_t modu(int i) { constexpr uint32_t t[] = PREFIX##_MODU; return t[i]; } \
        static ZKM_DEV uint32_t mod(int i) { return PREFIX##_MOD[i]; }              \
        static ZKM_DEV uint32_t one(int i) { return PREFIX##_ONE[i]; }              \
        static ZKM_DEV uint32_t r2(int i) { return PREFIX##_R2[i]; }                \
        static ZKM_DEV uint32_t r3(int i) { return PREFIX##_R3[i]; }                \
    };

ZKM_DEFINE_FP_PARAMS(Bls12_381_FqP, BLS12_381_FQ, 12)
ZKM_DEFINE_FP_PARAMS(Bls12_381_FrP, BLS12_381_FR, 8)
ZKM_DEFINE_FP_PARAMS(Bn254_FqP, BN254_FQ, 8)
ZKM_DEFINE_FP_PARAMS(Bn254_FrP, BN254_FR, 8)
ZKM_DEFINE_FP_PARAMS(Bw6_761_FqP, BW6_761_FQ, 24)
ZKM_DEFINE_FP_PARAMS(Bw6_761_FrP, BW6_761_FR, 12)   // = the base field of BLS12-377

// Low limbs of the modulus that make a reduction step free of multiplications (round 2): p_0 = 1 (p = 1 mod 2^32: the
// NTT-friendly scalar fields) turns m * p_0 into an addition of m; p_1 = 2^32 - 1 (BLS12-381 Fr) turns m * p_1 into
// (m - [m != 0]) : (-m), two ALU operations.  fp_mul_cios then issues N^2 + N (N - 2) instead of 2 N^2 wide MADs:
// 112 instead of 128 for the 8-limb field of every NTT and witness map.
template <class P> struct FpLowLimbs { static constexpr bool p0_one = false, p1_ones = false; };
template <> struct FpLowLimbs<Bls12_381_FrP> { static constexpr bool p0_one = true, p1_ones = true; };
template <> struct FpLowLimbs<Bw6_761_FrP> { static constexpr bool p0_one = true, p1_ones = false; };

// ----------------------------------------------------------------------------- Fp
template <class P>
struct Fp {
    static constexpr int N = P::N;
    typedef P Params;
    uint32_t l[N];

    static ZKM_DEV Fp zero() {
        Fp r;
        ZKM_UNROLL
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    static ZKM_DEV Fp one() {
        Fp r;
        ZKM_UNROLL
        for (int i = 0; i < N; i++) r.l[i] = P::one(i);
        return r;
    }
    static ZKM_DEV Fp r2() {
        Fp r;
        ZKM_UNROLL
        for (int i = 0; i < N; i++) r.l[i] = P::r2(i);
        return r;
    }
    ZKM_DEV bool is_zero() const {
        uint32_t o = 0;
        ZKM_UNROLL
        for (int i = 0; i < N; i++) o |= l[i];
        return o == 0;
    }
    ZKM_DEV bool operator==(const Fp& b) const {
        uint32_t o = 0;
        ZKM_UNROLL
        for (int i = 0; i < N; i++) o |= (l[i] ^ b.l[i]);
        return o == 0;
    }
    ZKM_DEV bool operator!=(const Fp& b) const { return !(*this == b); }
};

// r = (t >= p) ? t - p : t      (t < 2p)
template <class P>
ZKM_DEV void fp_final_sub(Fp<P>& t) {
    constexpr int N = P::N;
    uint32_t d[N];
    d[0] = ptx::sub_cc(t.l[0], P::mod(0));
    ZKM_UNROLL
    for (int i = 1; i < N; i++) d[i] = ptx::subc_cc(t.l[i], P::mod(i));
    uint32_t borrow = ptx::subc(0, 0);  // 0xffffffff when t < p
    ZKM_UNROLL
    for (int i = 0; i < N; i++) t.l[i] = borrow ? t.l[i] : d[i];
}

template <class P>
ZKM_DEV Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> r;
    r.l[0] = ptx::add_cc(a.l[0], b.l[0]);
    ZKM_UNROLL
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
    r.l[N - 1] = ptx::addc(a.l[N - 1], b.l[N - 1]);
    fp_final_sub(r);
    return r;
}

template <class P>
ZKM_DEV Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> r;
    r.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
    ZKM_UNROLL
    for (int i = 1; i < N; i++) r.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
    uint32_t mask = ptx::subc(0, 0);  // all ones when a < b
    r.l[0] = ptx::add_cc(r.l[0], P::mod(0) & mask);
    ZKM_UNROLL
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(r.l[i], P::mod(i) & mask);
    r.l[N - 1] = ptx::addc(r.l[N - 1], P::mod(N - 1) & mask);
    return r;
}

template <class P>
ZKM_DEV Fp<P> fp_neg(const Fp<P>& a) {
    constexpr int N = P::N;
    Fp<P> r;
    uint32_t nz = 0;
    ZKM_UNROLL
    for (int i = 0; i < N; i++) nz |= a.l[i];
    uint32_t mask = nz ? 0xffffffffu : 0u;  // -0 = 0
    r.l[0] = ptx::sub_cc(P::mod(0) & mask, a.l[0]);
    ZKM_UNROLL
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::subc_cc(P::mod(i) & mask, a.l[i]);
    r.l[N - 1] = ptx::subc(P::mod(N - 1) & mask, a.l[N - 1]);
    return r;
}

template <class P>
ZKM_DEV Fp<P> fp_dbl(const Fp<P>& a) {
    return fp_add(a, a);
}

// Montgomery product a*b*R^-1 mod p, fully reduced: the saturated 32-bit CIOS described above.  This is the
// product the kernels use (fp_mul below).  Round 2 measured the alternative -- unsaturated radix 2^30 columns of
// plain IMAD.WIDE.U32 without carry predicates (zkm_fpmul_u.cuh, fp_mul_unsat) -- and found it 0.65-0.9x: on
// sm_100a EVERY 32x32->64 product issues at ~31 /clk/SM whatever its form (IMAD.WIDE with or without a 64-bit
// addend, .X carry, IMAD.HI; tools/microbench/int_pipe_peak.cu, profiles/int_pipe_peak_r2.jsonl), so the
// carry-free schedule only adds ALU instructions.  Both products return identical bytes (host-emulation tests,
// tools/microbench/fpmul_bench.cu checks 2^20 pairs per field on the device).
template <class P>
ZKM_DEV Fp<P> fp_mul_cios(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    static_assert((N & 1) == 0, "even limb count required");
    // The reduction factor is formed with INV read from constant memory (P::inv_rt()), not the immediate: for the
    // fields with p = 1 mod 2^32 (INV = 0xffffffff: BLS12-381 Fr, BW6-761 Fr) ptxas otherwise rewrites m = -t and then
    // leaves every m * p_i pair unfused (IMAD.X + IMAD.HI.U32.X, 6 issue cycles instead of the 4 of one
    // IMAD.WIDE.U32.X): 280 instead of 228 instructions per 8-limb product (SASS: profiles/int_pipe_peak_r2.sass.txt).
    uint32_t acc[2][2 * N + 2];
    ZKM_UNROLL
    for (int i = 0; i < 2 * N + 2; i++) {
        acc[0][i] = 0;
        acc[1][i] = 0;
    }
    ZKM_UNROLL
    for (int j = 0; j < N; j++) {
        uint32_t* X = acc[j & 1];        // chain of even a-limbs starts at column j
        uint32_t* Y = acc[(j & 1) ^ 1];  // chain of odd a-limbs starts at column j+1
        const uint32_t bj = b.l[j];
        if (j == 0) {
            ZKM_UNROLL
            for (int i = 1; i < N; i += 2) {
                Y[i] = ptx::mul_lo(a.l[i], bj);
                Y[i + 1] = ptx::mul_hi(a.l[i], bj);
            }
            ZKM_UNROLL
            for (int i = 0; i < N; i += 2) {
                X[i] = ptx::mul_lo(a.l[i], bj);
                X[i + 1] = ptx::mul_hi(a.l[i], bj);
            }
        } else {
            // merge column j of the two accumulators; its carry enters the Y chain
            X[j] = ptx::add_cc(X[j], Y[j]);
            ZKM_UNROLL
            for (int i = 1; i < N; i += 2) {
                Y[j + i] = ptx::madc_lo_cc(a.l[i], bj, Y[j + i]);
                Y[j + i + 1] = ptx::madc_hi_cc(a.l[i], bj, Y[j + i + 1]);
            }
            Y[j + N + 1] = ptx::addc(Y[j + N + 1], 0);
            X[j] = ptx::mad_lo_cc(a.l[0], bj, X[j]);
            X[j + 1] = ptx::madc_hi_cc(a.l[0], bj, X[j + 1]);
            ZKM_UNROLL
            for (int i = 2; i < N; i += 2) {
                X[j + i] = ptx::madc_lo_cc(a.l[i], bj, X[j + i]);
                X[j + i + 1] = ptx::madc_hi_cc(a.l[i], bj, X[j + i + 1]);
            }
            X[j + N] = ptx::addc(X[j + N], 0);
        }
        const uint32_t m = ptx::mul_lo(X[j], P::inv_rt());
        if (FpLowLimbs<P>::p1_ones) {
            // m * (2^32 - 1) = (m - [m != 0]) * 2^32 + (2^32 - m) mod 2^32: no multiplication
            const uint32_t lo1 = ptx::sub_cc(0u, m);
            const uint32_t hi1 = ptx::subc(m, 0u);
            Y[j + 1] = ptx::add_cc(Y[j + 1], lo1);
            Y[j + 2] = ptx::addc_cc(Y[j + 2], hi1);
        } else {
            Y[j + 1] = ptx::mad_lo_cc(m, P::mod(1), Y[j + 1]);
            Y[j + 2] = ptx::madc_hi_cc(m, P::mod(1), Y[j + 2]);
        }
        ZKM_UNROLL
        for (int i = 3; i < N; i += 2) {
            Y[j + i] = ptx::madc_lo_cc(m, P::mod(i), Y[j + i]);
            Y[j + i + 1] = ptx::madc_hi_cc(m, P::mod(i), Y[j + i + 1]);
        }
        Y[j + N + 1] = ptx::addc(Y[j + N + 1], 0);
        if (FpLowLimbs<P>::p0_one) {
            X[j] = ptx::add_cc(X[j], m);              // m * 1: column j becomes 0, the carry is [m != 0]
            X[j + 1] = ptx::addc_cc(X[j + 1], 0u);
        } else {
            X[j] = ptx::mad_lo_cc(m, P::mod(0), X[j]);
            X[j + 1] = ptx::madc_hi_cc(m, P::mod(0), X[j + 1]);
        }
        ZKM_UNROLL
        for (int i = 2; i < N; i += 2) {
            X[j + i] = ptx::madc_lo_cc(m, P::mod(i), X[j + i]);
            X[j + i + 1] = ptx::madc_hi_cc(m, P::mod(i), X[j + i + 1]);
        }
        X[j + N] = ptx::addc(X[j + N], 0);
    }
    Fp<P> r;
    r.l[0] = ptx::add_cc(acc[0][N], acc[1][N]);
    ZKM_UNROLL
    for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(acc[0][N + i], acc[1][N + i]);
    r.l[N - 1] = ptx::addc(acc[0][2 * N - 1], acc[1][2 * N - 1]);
    fp_final_sub(r);
    return r;
}

// Montgomery product on the unsaturated radix (zkm_fpmul_u.cuh): plain IMAD.WIDE products, carries per column.
// Measured slower than the CIOS on B200 (see above); kept for the microbenchmark and as the cross-check.
template <class P>
ZKM_DEV Fp<P> fp_mul_unsat(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> r;
    fp_mul_u_raw<P>(r.l, fp_split<P>(a.l), fp_split<P>(b.l));
    fp_final_sub(r);
    return r;
}
// dedicated squaring on the same columns: UM (UM + 1) / 2 products for the a^2 half instead of UM^2
template <class P>
ZKM_DEV Fp<P> fp_sqr_unsat(const Fp<P>& a) {
    Fp<P> r;
    fp_sqr_u_raw<P>(r.l, fp_split<P>(a.l));
    fp_final_sub(r);
    return r;
}

template <class P>
ZKM_DEV Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
#if defined(ZKM_FPMUL_UNSAT)
    return fp_mul_unsat(a, b);
#else
    return fp_mul_cios(a, b);
#endif
}

template <class P>
ZKM_DEV Fp<P> fp_sqr(const Fp<P>& a) {
#if defined(ZKM_FPMUL_UNSAT)
    return fp_sqr_unsat(a);
#else
    return fp_mul_cios(a, a);
#endif
}

// a^e for a little-endian multi-limb exponent (setup / normalisation only).
template <class P>
ZKM_DEV Fp<P> fp_pow_limbs(const Fp<P>& a, const uint32_t* e, int nlimbs) {
    Fp<P> r = Fp<P>::one();
    bool started = false;
    for (int i = nlimbs - 1; i >= 0; i--) {
        for (int b = 31; b >= 0; b--) {
            if (started) r = fp_sqr(r);
            if ((e[i] >> b) & 1) {
                r = started ? fp_mul(r, a) : a;
                started = true;
            }
        }
    }
    return r;
}

template <class P>
ZKM_DEV Fp<P> fp_pow_u64(const Fp<P>& a, uint64_t e) {
    uint32_t w[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
    return fp_pow_limbs(a, w, 2);
}

// Fermat inverse a^(p-2) (kept as the cross-check of fp_inv in the host-emulation tests).
template <class P>
ZKM_DEV Fp<P> fp_inv_fermat(const Fp<P>& a) {
    constexpr int N = P::N;
    uint32_t e[N];
    uint32_t borrow = 2;
    for (int i = 0; i < N; i++) {
        uint32_t m = P::mod(i);
        e[i] = m - borrow;
        borrow = (m < borrow) ? 1u : 0u;
    }
    return fp_pow_limbs(a, e, N);
}

// ---- raw multi-limb helpers for the binary extended Euclid below (values < 2^(32 N))
template <class P>
ZKM_DEV void limbs_shr1(uint32_t (&x)[P::N]) {
    ZKM_UNROLL
    for (int i = 0; i < P::N - 1; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
    x[P::N - 1] >>= 1;
}
template <class P>
ZKM_DEV void limbs_halve_mod(uint32_t (&x)[P::N]) {  // x / 2 mod p for x < p  (p odd, spare top bit)
    uint32_t mask = (x[0] & 1u) ? 0xffffffffu : 0u;
    x[0] = ptx::add_cc(x[0], P::mod(0) & mask);
    ZKM_UNROLL
    for (int i = 1; i < P::N - 1; i++) x[i] = ptx::addc_cc(x[i], P::mod(i) & mask);
    x[P::N - 1] = ptx::addc(x[P::N - 1], P::mod(P::N - 1) & mask);
    limbs_shr1<P>(x);
}
template <class P>
ZKM_DEV uint32_t limbs_sub(uint32_t (&r)[P::N], const uint32_t (&a)[P::N], const uint32_t (&b)[P::N]) {  // returns borrow mask
    r[0] = ptx::sub_cc(a[0], b[0]);
    ZKM_UNROLL
    for (int i = 1; i < P::N; i++) r[i] = ptx::subc_cc(a[i], b[i]);
    return ptx::subc(0, 0);
}
template <class P>
ZKM_DEV bool limbs_is_one(const uint32_t (&x)[P::N]) {
    uint32_t o = x[0] ^ 1u;
    ZKM_UNROLL
    for (int i = 1; i < P::N; i++) o |= x[i];
    return o == 0;
}

// Modular inverse by the binary extended Euclidean algorithm on the Montgomery representative
// (x = (aR)^-1 mod p), then one Montgomery product with R^3 gives a^-1 R.  ~2 log2(p) shift/subtract
// steps of a few dozen instructions instead of ~1.5 log2(p) dependent Montgomery products: an order of
// magnitude less latency for the single-thread tails (final normalisation, batched-inversion roots).
// The value is the same as the Fermat inverse (the inverse is unique); inv(0) = 0 like ark-ff's None -> zero use.
template <class P>
ZKM_DEV Fp<P> fp_inv(const Fp<P>& a) {
    constexpr int N = P::N;
    if (a.is_zero()) return a;
    uint32_t u[N], v[N], t[N];
    Fp<P> x1, x2;
    ZKM_UNROLL
    for (int i = 0; i < N; i++) {
        u[i] = a.l[i];
        v[i] = P::mod(i);
        x1.l[i] = (i == 0) ? 1u : 0u;
        x2.l[i] = 0u;
    }
    while (!limbs_is_one<P>(u) && !limbs_is_one<P>(v)) {
        while ((u[0] & 1u) == 0) {
            limbs_shr1<P>(u);
            limbs_halve_mod<P>(x1.l);
        }
        while ((v[0] & 1u) == 0) {
            limbs_shr1<P>(v);
            limbs_halve_mod<P>(x2.l);
        }
        uint32_t borrow = limbs_sub<P>(t, u, v);
        if (borrow == 0) {  // u >= v
            ZKM_UNROLL
            for (int i = 0; i < N; i++) u[i] = t[i];
            x1 = fp_sub(x1, x2);
        } else {
            limbs_sub<P>(v, v, u);
            x2 = fp_sub(x2, x1);
        }
    }
    Fp<P> x = limbs_is_one<P>(u) ? x1 : x2;
    Fp<P> r3;
    ZKM_UNROLL
    for (int i = 0; i < N; i++) r3.l[i] = P::r3(i);
    return fp_mul(x, r3);
}

template <class P> ZKM_DEV Fp<P> operator+(const Fp<P>& a, const Fp<P>& b) { return fp_add(a, b); }
template <class P> ZKM_DEV Fp<P> operator-(const Fp<P>& a, const Fp<P>& b) { return fp_sub(a, b); }
template <class P> ZKM_DEV Fp<P> operator*(const Fp<P>& a, const Fp<P>& b) { return fp_mul(a, b); }
template <class P> ZKM_DEV Fp<P> sqr(const Fp<P>& a) { return fp_sqr(a); }
template <class P> ZKM_DEV Fp<P> dbl(const Fp<P>& a) { return fp_dbl(a); }
template <class P> ZKM_DEV Fp<P> neg(const Fp<P>& a) { return fp_neg(a); }
template <class P> ZKM_DEV Fp<P> inv(const Fp<P>& a) { return fp_inv(a); }

// ----------------------------------------------------------------------------- Fp2 = Fp[u]/(u^2+1)
template <class P>
struct Fp2 {
    typedef P Params;
    Fp<P> c0, c1;
    static ZKM_DEV Fp2 zero() { Fp2 r; r.c0 = Fp<P>::zero(); r.c1 = Fp<P>::zero(); return r; }
    static ZKM_DEV Fp2 one() { Fp2 r; r.c0 = Fp<P>::one(); r.c1 = Fp<P>::zero(); return r; }
    ZKM_DEV bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    ZKM_DEV bool operator==(const Fp2& b) const { return c0 == b.c0 && c1 == b.c1; }
    ZKM_DEV bool operator!=(const Fp2& b) const { return !(*this == b); }
};

template <class P> ZKM_DEV Fp2<P> operator+(const Fp2<P>& a, const Fp2<P>& b) { Fp2<P> r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
template <class P> ZKM_DEV Fp2<P> operator-(const Fp2<P>& a, const Fp2<P>& b) { Fp2<P> r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
template <class P> ZKM_DEV Fp2<P> neg(const Fp2<P>& a) { Fp2<P> r; r.c0 = neg(a.c0); r.c1 = neg(a.c1); return r; }
template <class P> ZKM_DEV Fp2<P> dbl(const Fp2<P>& a) { Fp2<P> r; r.c0 = dbl(a.c0); r.c1 = dbl(a.c1); return r; }
// Karatsuba: (a0 b0 - a1 b1) + ((a0+a1)(b0+b1) - a0 b0 - a1 b1) u
template <class P>
ZKM_DEV Fp2<P> operator*(const Fp2<P>& a, const Fp2<P>& b) {
    Fp<P> v0 = a.c0 * b.c0;
    Fp<P> v1 = a.c1 * b.c1;
    Fp<P> s = (a.c0 + a.c1) * (b.c0 + b.c1);
    Fp2<P> r;
    r.c0 = v0 - v1;
    r.c1 = (s - v0) - v1;
    return r;
}
// (a0+a1)(a0-a1) + 2 a0 a1 u
template <class P>
ZKM_DEV Fp2<P> sqr(const Fp2<P>& a) {
    Fp<P> t = a.c0 * a.c1;
    Fp2<P> r;
    r.c0 = (a.c0 + a.c1) * (a.c0 - a.c1);
    r.c1 = dbl(t);
    return r;
}
template <class P>
ZKM_DEV Fp2<P> inv(const Fp2<P>& a) {
    Fp<P> n = inv(sqr(a.c0) + sqr(a.c1));
    Fp2<P> r;
    r.c0 = a.c0 * n;
    r.c1 = neg(a.c1 * n);
    return r;
}

typedef Fp<Bls12_381_FqP> Bls12_381_Fq;
typedef Fp<Bls12_381_FrP> Bls12_381_Fr;
typedef Fp<Bn254_FqP> Bn254_Fq;
typedef Fp<Bn254_FrP> Bn254_Fr;
typedef Fp2<Bls12_381_FqP> Bls12_381_Fq2;
typedef Fp2<Bn254_FqP> Bn254_Fq2;
typedef Fp<Bw6_761_FqP> Bw6_761_Fq;   // G1 AND G2 of BW6-761 live over this 761-bit prime field
typedef Fp<Bw6_761_FrP> Bw6_761_Fr;

}  // namespace zkm

// zkm_ntt.cuh -- radix-2 NTT of the scalar field Fr on sm_100a.
//
// Replaces ark-poly 0.3.0 Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place
// (src/domain/radix2/{mod,fft}.rs: in_order_fft_in_place = io_helper + derange; ifft =
// derange + oi_helper + size_inv; src/domain/mod.rs: distribute_powers for the coset variants;
// pin /root/reference/Cargo.lock:338-339; reached from /root/reference/benches/groth16.rs:115 via
// ark-groth16's witness_map and from benches/marlin.rs:202,311 via the AHP prover).
// Outputs are unique field elements, so any correct evaluation order is byte-identical.
//
// Design (not a translation of io_helper/oi_helper): the log2(n) decimation-in-frequency stages
// are grouped into passes of R <= 12 stages.  A pass is the "four-step" split of the remaining
// sub-transform of size n' = 2^(k - s0):  every tile gathers 2^R elements with stride
// L = n' / 2^R, runs a 2^R-point DIF NTT on chip (8 elements per thread in registers, 3 stages
// per round, rounds exchanged through XOR-swizzled shared memory), multiplies output u of the
// tile by w_{n'}^(lo * u) (one table look-up) and writes it back in place at the bit-reversed
// slot.  The last pass writes each value straight to its natural-order position (the global
// bit reversal of `derange` folded into the store), so it is out of place.  Coset scaling
// (x_j * g^j on the first load; g^-j * n^-1 on the last store) and size_inv are fused in.
// Twiddles are precomputed tables kept in HBM/L2 and reused by every later call.
//
// Cost per element: R/2 butterflies per pass (the last stage of a pass has unit twiddles and
// skips its multiply) + 1 four-step multiply per non-final pass.  HBM traffic: 64 B per element
// per pass.  On B200 the kernel is bound by the integer pipe (Montgomery products), not by HBM:
// see DESIGN.md.
#pragma once
#include "zkm_common.cuh"

namespace zkm {

template <class P> struct FrRoots;
template <> struct FrRoots<Bls12_381_FrP> {
    static constexpr int TWO_ADICITY = BLS12_381_FR_TWO_ADICITY;
    static __device__ uint32_t root(int i) { return BLS12_381_FR_ROOT[i]; }
    static __device__ uint32_t root_inv(int i) { return BLS12_381_FR_ROOT_INV[i]; }
    static __device__ uint32_t gen(int i) { return BLS12_381_FR_GEN[i]; }
    static __device__ uint32_t gen_inv(int i) { return BLS12_381_FR_GEN_INV[i]; }
    static __device__ uint32_t two_inv(int i) { return BLS12_381_FR_TWO_INV[i]; }
};
template <> struct FrRoots<Bn254_FrP> {
    static constexpr int TWO_ADICITY = BN254_FR_TWO_ADICITY;
    static __device__ uint32_t root(int i) { return BN254_FR_ROOT[i]; }
    static __device__ uint32_t root_inv(int i) { return BN254_FR_ROOT_INV[i]; }
    static __device__ uint32_t gen(int i) { return BN254_FR_GEN[i]; }
    static __device__ uint32_t gen_inv(int i) { return BN254_FR_GEN_INV[i]; }
    static __device__ uint32_t two_inv(int i) { return BN254_FR_TWO_INV[i]; }
};

template <> struct FrRoots<Bw6_761_FrP> {   // 377-bit Fr of BW6-761 = Fq of BLS12-377: GENERATOR = -5, two-adicity 46
    static constexpr int TWO_ADICITY = BW6_761_FR_TWO_ADICITY;
    static __device__ uint32_t root(int i) { return BW6_761_FR_ROOT[i]; }
    static __device__ uint32_t root_inv(int i) { return BW6_761_FR_ROOT_INV[i]; }
    static __device__ uint32_t gen(int i) { return BW6_761_FR_GEN[i]; }
    static __device__ uint32_t gen_inv(int i) { return BW6_761_FR_GEN_INV[i]; }
    static __device__ uint32_t two_inv(int i) { return BW6_761_FR_TWO_INV[i]; }
};

template <class P>
__device__ Fp<P> fr_const(uint32_t (*f)(int)) {
    Fp<P> r;
    for (int i = 0; i < P::N; i++) r.l[i] = f(i);
    return r;
}

// w_k = ROOT^(2^(adicity - k))  (or its inverse): generator of the order-2^k subgroup
template <class P>
__device__ Fp<P> group_gen(int k, bool inverse) {
    Fp<P> w = inverse ? fr_const<P>(FrRoots<P>::root_inv) : fr_const<P>(FrRoots<P>::root);
    for (int i = k; i < FrRoots<P>::TWO_ADICITY; i++) w = fp_sqr(w);
    return w;
}

// ---------------------------------------------------------------------------------- table generation
// out[e] = w_k^e (or w_k^-e), e < 2^(k-1)
template <class P>
__global__ void k_gen_twiddles(uint32_t* out, int k, int inverse, uint64_t count) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    Fp<P> w = group_gen<P>(k, inverse != 0);
    Fp<P> r = fp_pow_u64(w, e);
    st_fp<P>(out + e * P::N, r);
}

// four-step twiddles of a non-final pass in TILE order: out[(lo << R) + j] = w_kk^(lo * bitrev_R(j)), lo < 2^(kk - R)
template <class P>
__global__ void k_gen_tw4(uint32_t* out, int kk, int R, int inverse) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (1ull << kk)) return;
    const uint64_t lo = idx >> R;
    const uint32_t j = (uint32_t)(idx & ((1u << R) - 1u));
    const uint32_t u = __brev(j) >> (32 - R);
    Fp<P> w = group_gen<P>(kk, inverse != 0);
    st_fp<P>(out + idx * P::N, fp_pow_u64(w, lo * (uint64_t)u));
}

// coset tables: lo[j] = c * g^(+-j), j < nlo ; hi[j] = g^(+-j*nlo), j < nhi ; c = 1 (forward) or n^-1 (inverse)
template <class P>
__global__ void k_gen_coset(uint32_t* lo, uint32_t* hi, uint64_t nlo, uint64_t nhi, int inverse, int k) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nlo + nhi) return;
    Fp<P> g = inverse ? fr_const<P>(FrRoots<P>::gen_inv) : fr_const<P>(FrRoots<P>::gen);
    if (t < nlo) {
        Fp<P> r = fp_pow_u64(g, t);
        if (inverse) {
            Fp<P> ninv = fp_pow_u64(fr_const<P>(FrRoots<P>::two_inv), (uint64_t)k);
            r = r * ninv;
        }
        st_fp<P>(lo + t * P::N, r);
    } else {
        uint64_t j = t - nlo;
        Fp<P> r = fp_pow_u64(g, j * nlo);
        st_fp<P>(hi + j * P::N, r);
    }
}

// group_gen, group_gen_inv, size_inv, generator, generator_inv
template <class P>
__global__ void k_domain_constants(uint32_t* out, int k) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    st_fp<P>(out + 0 * P::N, group_gen<P>(k, false));
    st_fp<P>(out + 1 * P::N, group_gen<P>(k, true));
    st_fp<P>(out + 2 * P::N, fp_pow_u64(fr_const<P>(FrRoots<P>::two_inv), (uint64_t)k));
    st_fp<P>(out + 3 * P::N, fr_const<P>(FrRoots<P>::gen));
    st_fp<P>(out + 4 * P::N, fr_const<P>(FrRoots<P>::gen_inv));
}

// ---------------------------------------------------------------------------------- the pass kernel
struct NttPassArgs {
    const uint32_t* in;
    uint32_t* out;
    const uint32_t* tw_tile;   // w_{2^R}^e, e < 2^(R-1)
    const uint32_t* tw_four;   // w_{n'}^e, e < n'/2 (non-final passes; fallback when the per-tile table would be too large)
    const uint32_t* tw_four_tile;  // w_{n'}^(lo * bitrev_R(j)) laid out [lo][j]: what tile `lo` multiplies its slot j by,
                               // in the order its threads hold the slots -> one coalesced 2^R-element read per tile
    const uint32_t* coset_lo;  // fused coset / size_inv scaling (first or last pass), may be null
    const uint32_t* coset_hi;
    const uint32_t* size_inv;  // n^-1 (non-coset inverse, last pass)
    int k;                     // log2 n
    int s0;                    // stages already done
    int first, last;           // pass position
    int scale_in;              // multiply inputs by coset table (forward coset, first pass)
    int scale_out;             // 1: multiply outputs by coset table ; 2: by size_inv (last pass)
    int coset_lo_bits;
};

// NL = 32-bit limbs per element: 8 (255/254-bit Fr) or 12 (377-bit Fr of BW6-761)
template <int R, int NL>
struct NttCfg {
    static constexpr int TPT = 1 << (R - 3);               // threads per tile
    static constexpr int NT = TPT > 128 ? TPT : 128;       // threads per CTA
    static constexpr int TILES = NT / TPT;                 // tiles per CTA iteration
    static constexpr int SMEM = NT * 8 * NL * 4;           // bytes
    static constexpr int NR = (R + 2) / 3;                 // register rounds
    // 8 limbs: 16 warps per SM (<= 128 registers); 8 warps/SM measured 10 % slower.  12 limbs: the eight
    // register-resident elements alone are 96 registers, so 8 warps per SM with up to 255 registers (R <= 11).
    static constexpr int MINB = NL <= 8 ? 512 / NT : (256 / NT > 0 ? 256 / NT : 1);
};
template <class P> struct NttMaxR { static constexpr int value = P::N <= 8 ? 12 : 11; };

__device__ __forceinline__ uint32_t ntt_slot(uint32_t tau, uint32_t q, int pl) {
    return ((tau >> pl) << (pl + 3)) | (q << pl) | (tau & ((1u << pl) - 1u));
}
__device__ __forceinline__ uint32_t ntt_phys(uint32_t slot) { return slot ^ ((slot >> 3) & 7u); }

// Register rotation q -> rotl3(q): after it, the bit the NEXT stage butterflies on sits in position 2,
// so every stage runs the same four butterflies (i, i + 4).  One loop body of four Montgomery products
// instead of twelve unrolled ones keeps the round inside the instruction cache.
template <class P>
__device__ __forceinline__ void ntt_rotate(Fp<P> (&x)[8]) {
    Fp<P> t = x[4];
    x[4] = x[2];
    x[2] = x[1];
    x[1] = t;
    t = x[5];
    x[5] = x[6];
    x[6] = x[3];
    x[3] = t;
}

// `ns` DIF stages (q-bits ns-1 .. 0 of the 8 register-resident elements, identity mapping x[q] <-> q on
// entry and on exit)
template <class P, int R>
__device__ __forceinline__ void ntt_round(Fp<P> (&x)[8], uint32_t tau, int pl, int ns, const uint32_t* tw_tile) {
    int rot = (3 - ns) % 3;  // rotations applied so far: register i holds original q = rotr3(i, rot)
    if (rot >= 1) ntt_rotate<P>(x);
    if (rot == 2) ntt_rotate<P>(x);
#pragma unroll 1
    for (int s = 0; s < ns; s++) {
        const int beta = 2 - rot;      // original q-bit of this stage
        const int g = pl + beta;       // bit position of the butterfly distance inside the tile
        const int t = R - 1 - g;       // DIF stage index
#pragma unroll
        for (int i = 0; i < 4; i++) {
            Fp<P> lo = x[i], hi = x[i + 4];
            x[i] = fp_add(lo, hi);
            Fp<P> d = fp_sub(lo, hi);
            if (g == 0) {
                x[i + 4] = d;  // w^0
            } else {
                const uint32_t q = (((uint32_t)i >> rot) | ((uint32_t)i << (3 - rot))) & 7u;
                const uint32_t slot = ntt_slot(tau, q, pl);
                const uint32_t j = slot & ((1u << g) - 1u);
                if (j == 0) {
                    // w^0 as well.  In the last round of a pass (pl == 0) j depends on q only, so the branch is uniform
                    // over the CTA and 3 of its 8 remaining products disappear; elsewhere the few lanes with j == 0
                    // just sit out the product.
                    x[i + 4] = d;
                } else {
                    Fp<P> w = ld_fp<P>(tw_tile + ((size_t)j << t) * P::N);
                    x[i + 4] = fp_mul(d, w);
                }
            }
        }
        ntt_rotate<P>(x);
        rot++;
    }
}

template <class P>
__device__ __forceinline__ Fp<P> coset_factor(const NttPassArgs& a, uint64_t idx) {
    Fp<P> lo = ld_fp<P>(a.coset_lo + (idx & ((1ull << a.coset_lo_bits) - 1ull)) * P::N);
    Fp<P> hi = ld_fp<P>(a.coset_hi + (idx >> a.coset_lo_bits) * P::N);
    return fp_mul(lo, hi);
}

template <class P, int R>
__global__ void __launch_bounds__((NttCfg<R, P::N>::NT), (NttCfg<R, P::N>::MINB)) k_ntt_pass(const NttPassArgs a) {
    typedef NttCfg<R, P::N> C;
    constexpr int PLANES = P::N / 4;    // one shared-memory plane per 128-bit quarter of an element
    extern __shared__ uint4 ntt_smem[];
    const uint32_t tl = threadIdx.x / C::TPT;
    const uint32_t tau = threadIdx.x % C::TPT;
    uint4* sm0 = ntt_smem + (size_t)tl * ((uint32_t)PLANES << R);

    const int kk = a.k - a.s0;          // log2 n'
    const int logL = kk - R;            // log2 stride
    const uint64_t num_tiles = 1ull << (a.k - R);
    const uint64_t lmask = (1ull << logL) - 1ull;

    for (uint64_t base = (uint64_t)blockIdx.x * C::TILES; base < num_tiles; base += (uint64_t)gridDim.x * C::TILES) {
        const uint64_t tile = base + tl;
        const bool active = tile < num_tiles;
        const uint64_t lo_idx = tile & lmask;
        const uint64_t hi_idx = tile >> logL;
        const uint64_t gbase = (hi_idx << kk) + lo_idx;
        Fp<P> x[8];
        if (active) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                uint64_t gi = gbase + ((uint64_t)ntt_slot(tau, q, R - 3) << logL);
                x[q] = ld_fp_plain<P>(a.in + gi * P::N);
                if (a.scale_in) x[q] = fp_mul(x[q], coset_factor<P>(a, gi));
            }
        }
#pragma unroll 1
        for (int r = 0; r < C::NR; r++) {
            const int rem = R - 3 * r;
            const int ns = rem >= 3 ? 3 : rem;
            const int pl = rem >= 3 ? rem - 3 : 0;
            if (r > 0) {
                const int plp = R - 3 * r;  // previous round's pl (>= 0)
                __syncthreads();
                if (active) {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        uint32_t ph = ntt_phys(ntt_slot(tau, q, plp));
#pragma unroll
                        for (int v = 0; v < PLANES; v++)
                            sm0[((uint32_t)v << R) + ph] =
                                make_uint4(x[q].l[4 * v], x[q].l[4 * v + 1], x[q].l[4 * v + 2], x[q].l[4 * v + 3]);
                    }
                }
                __syncthreads();
                if (active) {
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        uint32_t ph = ntt_phys(ntt_slot(tau, q, pl));
#pragma unroll
                        for (int v = 0; v < PLANES; v++) {
                            uint4 t4 = sm0[((uint32_t)v << R) + ph];
                            x[q].l[4 * v] = t4.x; x[q].l[4 * v + 1] = t4.y; x[q].l[4 * v + 2] = t4.z; x[q].l[4 * v + 3] = t4.w;
                        }
                    }
                }
            }
            if (active) {
                ntt_round<P, R>(x, tau, pl, ns, a.tw_tile);
            }
        }
        if (active) {
            // the thread now holds slots j = 8 tau + q, i.e. outputs u = bitrev_R(j)
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint32_t j = (tau << 3) | q;
                const uint32_t u = __brev(j) >> (32 - R);
                if (!a.last) {
                    // four-step twiddle w_{n'}^(lo * u), then in place at the bit-reversed slot
                    Fp<P> v;
                    if (a.tw_four_tile) {
                        v = fp_mul(x[q], ld_fp<P>(a.tw_four_tile + ((lo_idx << R) + j) * P::N));
                    } else {
                        uint64_t e = lo_idx * (uint64_t)u;
                        const uint64_t half = 1ull << (kk - 1);
                        bool negate = e >= half;
                        if (negate) e -= half;
                        Fp<P> w = ld_fp<P>(a.tw_four + e * P::N);
                        v = fp_mul(x[q], w);
                        if (negate) v = fp_neg(v);
                    }
                    st_fp<P>(a.out + (gbase + ((uint64_t)j << logL)) * P::N, v);
                } else {
                    // natural-order position: u * 2^(k-R) + bitrev_{k-R}(hi)
                    const int hb = a.k - R;
                    uint64_t hrev = hb ? (uint64_t)(__brevll(hi_idx) >> (64 - hb)) : 0ull;
                    uint64_t K = ((uint64_t)u << hb) | hrev;
                    Fp<P> v = x[q];
                    if (a.scale_out == 1) {
                        v = fp_mul(v, coset_factor<P>(a, K));
                    } else if (a.scale_out == 2) {
                        v = fp_mul(v, ld_fp<P>(a.size_inv));
                    }
                    st_fp<P>(a.out + K * P::N, v);
                }
            }
        }
    }
}

// n <= 4: direct evaluation of the definition by one thread (k = 0, 1, 2)
template <class P>
__global__ void k_ntt_tiny(const uint32_t* in, uint32_t* out, int k, int inverse, int coset) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int n = 1 << k;
    Fp<P> x[4], y[4];
    Fp<P> g = inverse ? fr_const<P>(FrRoots<P>::gen_inv) : fr_const<P>(FrRoots<P>::gen);
    for (int i = 0; i < n; i++) {
        x[i] = ld_fp_plain<P>(in + i * P::N);
        if (coset && !inverse) x[i] = x[i] * fp_pow_u64(g, (uint64_t)i);
    }
    Fp<P> w = group_gen<P>(k, inverse != 0);
    Fp<P> ninv = fp_pow_u64(fr_const<P>(FrRoots<P>::two_inv), (uint64_t)k);
    for (int o = 0; o < n; o++) {
        Fp<P> acc = Fp<P>::zero();
        for (int i = 0; i < n; i++) acc = acc + x[i] * fp_pow_u64(w, (uint64_t)((i * o) % n));
        if (inverse) {
            acc = acc * ninv;
            if (coset) acc = acc * fp_pow_u64(g, (uint64_t)o);
        }
        y[o] = acc;
    }
    for (int i = 0; i < n; i++) st_fp<P>(out + i * P::N, y[i]);
}

// ---------------------------------------------------------------------------------- witness map glue
// ab[i] = (a[i] * b[i] - c[i]) * zinv,  zinv = 1 / (g^n - 1): the pointwise step of ark-groth16 0.3.0
// R1CStoQAP::witness_map (src/r1cs_to_qap.rs) between the coset FFTs and the coset iFFT.
template <class P>
__global__ void __launch_bounds__(256) k_qap_pointwise(uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                       const uint32_t* __restrict__ c, const uint32_t* __restrict__ zinv,
                                                       uint64_t n) {
    const Fp<P> z = ld_fp<P>(zinv);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        Fp<P> x = ld_fp_plain<P>(a + i * P::N);
        Fp<P> y = ld_fp<P>(b + i * P::N);
        Fp<P> w = ld_fp<P>(c + i * P::N);
        st_fp<P>(a + i * P::N, fp_mul(fp_sub(fp_mul(x, y), w), z));
    }
}
// Fr::into_repr(): Montgomery -> canonical integer (the scalar format of multi_scalar_mul); device-resident
// glue between the witness map (Montgomery coefficients h) and the h-query MSM.
template <class P>
__global__ void __launch_bounds__(256) k_fr_into_repr(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, uint64_t n) {
    Fp<P> one_raw = Fp<P>::zero();
    one_raw.l[0] = 1;   // the integer 1: a * 1 * R^-1 = canonical a
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        st_fp<P>(out + i * P::N, fp_mul(ld_fp_plain<P>(in + i * P::N), one_raw));
}
// out[0] = 1 / (g^(2^k) - 1)
template <class P>
__global__ void k_vanishing_inv(uint32_t* out, int k) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    Fp<P> g = fr_const<P>(FrRoots<P>::gen);
    for (int i = 0; i < k; i++) g = fp_sqr(g);
    st_fp<P>(out, fp_inv(fp_sub(g, Fp<P>::one())));
}

// ---------------------------------------------------------------------------------- KZG10 witness polynomial
// q(X) = (p(X) - p(z)) / (X - z): the division KZG10::open performs before its MSM (ark-poly-commit 0.3.0
// src/kzg10/mod.rs compute_witness_polynomial = p / (X - point); reached from /root/reference/benches/marlin.rs:311
// through MarlinKZG10::open).  Synthetic division is the recurrence q_{i-1} = p_i + z q_i; here it is evaluated as
// a blocked scan: with H_c = sum_{j >= c L} p_j z^(j - c L) the value at every chunk boundary,
//   k_quot_chunk_sums   S_c = sum_{j < L} p_{cL + j} z^j                       one thread per chunk of L coefficients
//   k_quot_boundaries   H_c = S_c + z^L H_{c+1}  for all c                     one CTA: serial inside a thread's run of
//                                                                              chunks, Hillis-Steele suffix scan across threads
//   k_quot_fill         q_i, i in chunk c, by the recurrence started from H_{c+1}
// H_0 = p(z) comes for free (the evaluation KZG10::open needs of the blinding polynomial).  Exact field arithmetic:
// the quotient is unique, so the bytes equal upstream's DensePolynomial division.
constexpr int ZKM_QUOT_L = 32;
constexpr int ZKM_QUOT_T = 512;

template <class P>
__global__ void __launch_bounds__(256) k_quot_chunk_sums(const uint32_t* __restrict__ p, uint64_t n, const uint32_t* __restrict__ zp,
                                                         uint32_t* __restrict__ S, uint64_t nchunks) {
    const Fp<P> z = ld_fp<P>(zp);
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += (uint64_t)gridDim.x * blockDim.x) {
        Fp<P> acc = Fp<P>::zero();
        const uint64_t lo = c * ZKM_QUOT_L;
        for (int j = ZKM_QUOT_L - 1; j >= 0; j--) {
            acc = fp_mul(acc, z);
            if (lo + j < n) acc = fp_add(acc, ld_fp<P>(p + (lo + j) * P::N));
        }
        st_fp<P>(S + c * P::N, acc);
    }
}

// S (padded with zeros to ZKM_QUOT_T * per chunks) -> H[c] for c in [0, nchunks]; H[nchunks] = 0
template <class P>
__global__ void __launch_bounds__(ZKM_QUOT_T) k_quot_boundaries(const uint32_t* __restrict__ S, uint64_t nchunks, uint64_t per,
                                                                const uint32_t* __restrict__ zp, uint32_t* __restrict__ H) {
    extern __shared__ uint4 quot_smem[];
    uint32_t* A = reinterpret_cast<uint32_t*>(quot_smem);
    const uint32_t t = threadIdx.x;
    Fp<P> zL = ld_fp<P>(zp);
    for (int i = 1; i < ZKM_QUOT_L; i <<= 1) zL = fp_sqr(zL);            // z^L, L a power of two
    const uint64_t c0 = (uint64_t)t * per;
    // A_t = sum_{c in own run} S_c zL^(c - c0)
    Fp<P> acc = Fp<P>::zero();
    for (uint64_t i = per; i-- > 0;) {
        acc = fp_mul(acc, zL);
        if (c0 + i < nchunks) acc = fp_add(acc, ld_fp_plain<P>(S + (c0 + i) * P::N));
    }
    st_fp<P>(A + t * P::N, acc);
    Fp<P> pd = fp_pow_u64(zL, per);                                      // weight of one run; squared every step
    __syncthreads();
    for (uint32_t d = 1; d < ZKM_QUOT_T; d <<= 1) {
        Fp<P> other = Fp<P>::zero();
        const bool has = t + d < ZKM_QUOT_T;
        if (has) other = ld_fp_plain<P>(A + (t + d) * P::N);
        __syncthreads();
        if (has) {
            acc = fp_add(acc, fp_mul(pd, other));
            st_fp<P>(A + t * P::N, acc);
        }
        pd = fp_sqr(pd);
        __syncthreads();
    }
    // acc = H at the bottom of the own run; walk the run top-down from the run above
    Fp<P> v = (t + 1 < ZKM_QUOT_T) ? ld_fp_plain<P>(A + (t + 1) * P::N) : Fp<P>::zero();
    for (uint64_t i = per; i-- > 0;) {
        const uint64_t c = c0 + i;
        if (c + 1 <= nchunks) st_fp<P>(H + (c + 1) * P::N, v);          // v == H_{c+1} here
        Fp<P> sc = (c < nchunks) ? ld_fp_plain<P>(S + c * P::N) : Fp<P>::zero();
        v = fp_add(sc, fp_mul(zL, v));
    }
    if (t == 0) st_fp<P>(H, v);                                          // H_0 = p(z)
}

template <class P>
__global__ void __launch_bounds__(256) k_quot_fill(const uint32_t* __restrict__ p, uint64_t n, const uint32_t* __restrict__ zp,
                                                   const uint32_t* __restrict__ H, uint32_t* __restrict__ q, uint64_t nchunks) {
    const Fp<P> z = ld_fp<P>(zp);
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += (uint64_t)gridDim.x * blockDim.x) {
        Fp<P> v = ld_fp_plain<P>(H + (c + 1) * P::N);
        const uint64_t lo = c * ZKM_QUOT_L;
        for (int j = ZKM_QUOT_L - 1; j >= 0; j--) {
            const uint64_t i = lo + j;
            if (i >= n) continue;                                        // beyond the top coefficient: v stays 0
            if (i + 1 < n) st_fp<P>(q + i * P::N, v);                    // q has n - 1 coefficients
            v = fp_add(ld_fp<P>(p + i * P::N), fp_mul(z, v));
        }
    }
}

template <class P>
static void kzg_quotient_t(Context* c, const uint64_t* d_coeffs, size_t n, const uint64_t* d_point, uint64_t* d_quot,
                           uint64_t* d_eval, cudaStream_t s) {
    if (n == 0) {
        if (d_eval) ZKM_CUDA(cudaMemsetAsync(d_eval, 0, P::N * 4, s));
        return;
    }
    const uint64_t nchunks = (n + ZKM_QUOT_L - 1) / ZKM_QUOT_L;
    const uint64_t per = (nchunks + ZKM_QUOT_T - 1) / ZKM_QUOT_T;
    uint32_t* S = (uint32_t*)c->ntt_b.get((2 * nchunks + 2) * P::N * 4);
    uint32_t* H = S + (nchunks + 1) * P::N;
    uint64_t blocks = (nchunks + 255) / 256, cap = (uint64_t)c->sm_count * 8;
    const unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
    ZKM_LAUNCH(k_quot_chunk_sums<P>, grid, 256, 0, s, (const uint32_t*)d_coeffs, (uint64_t)n, (const uint32_t*)d_point, S, nchunks);
    ZKM_LAUNCH(k_quot_boundaries<P>, 1, ZKM_QUOT_T, ZKM_QUOT_T * P::N * 4, s, (const uint32_t*)S, nchunks, per,
               (const uint32_t*)d_point, H);
    ZKM_LAUNCH(k_quot_fill<P>, grid, 256, 0, s, (const uint32_t*)d_coeffs, (uint64_t)n, (const uint32_t*)d_point,
               (const uint32_t*)H, (uint32_t*)d_quot, nchunks);
    if (d_eval) ZKM_CUDA(cudaMemcpyAsync(d_eval, H, P::N * 4, cudaMemcpyDeviceToDevice, s));
}

// ---------------------------------------------------------------------------------- host side
enum TableKind : uint64_t { TW = 1, COSET_LO = 2, COSET_HI = 3, DOMAIN = 4, VANISH = 5, TW4 = 6 };
static uint64_t table_key(uint64_t kind, int curve, int k, int inverse) {
    return (kind << 32) | ((uint64_t)curve << 16) | ((uint64_t)inverse << 8) | (uint64_t)k;
}

template <class P>
static const uint32_t* get_twiddles(Context* c, int curve, int k, int inverse, cudaStream_t s) {
    if (k < 1) k = 1;
    std::lock_guard<std::mutex> tw_lock(c->sh->tw_mu);
    uint64_t key = table_key(TW, curve, k, inverse);
    auto it = c->twiddles.find(key);
    if (it != c->twiddles.end()) return (const uint32_t*)it->second;
    uint64_t count = 1ull << (k - 1);
    void* p = nullptr;
    ZKM_CUDA(malloc_retry((void**)&p, count * P::N * 4));
    c->twiddles[key] = p;
    unsigned blocks = (unsigned)((count + 255) / 256);
    ZKM_LAUNCH(k_gen_twiddles<P>, blocks, 256, 0, s, (uint32_t*)p, k, inverse, count);
    ZKM_CUDA(cudaStreamSynchronize(s));  // first use only: other lanes' streams may read the table right away
    return (const uint32_t*)p;
}

// Per-tile four-step table of a non-final pass (2^kk elements).  The gather table `tw_four` costs a random 32-byte
// DRAM access per element (ncu r1: 1.84 GB read in pass 1 of a 2^24 transform for 0.54 GB of data); this layout is
// read front to back.  Built once per (size, tile radix, direction); above ZKM_TW4_MAX_BYTES the gather table is used.
constexpr size_t ZKM_TW4_MAX_BYTES = (size_t)512 << 20;
template <class P>
static const uint32_t* get_tw4(Context* c, int curve, int kk, int R, int inverse, cudaStream_t s) {
    const size_t bytes = ((size_t)P::N * 4) << kk;
    if (bytes > ZKM_TW4_MAX_BYTES) return nullptr;
    std::lock_guard<std::mutex> tw_lock(c->sh->tw_mu);
    const uint64_t key = table_key(TW4, curve, kk, inverse) | ((uint64_t)R << 40);
    auto it = c->twiddles.find(key);
    if (it != c->twiddles.end()) return (const uint32_t*)it->second;
    void* p = nullptr;
    ZKM_CUDA(malloc_retry((void**)&p, bytes));
    c->twiddles[key] = p;
    ZKM_LAUNCH(k_gen_tw4<P>, (unsigned)(((1ull << kk) + 255) / 256), 256, 0, s, (uint32_t*)p, kk, R, inverse);
    ZKM_CUDA(cudaStreamSynchronize(s));
    return (const uint32_t*)p;
}

template <class P>
static void get_coset(Context* c, int curve, int k, int inverse, cudaStream_t s, const uint32_t** lo,
                      const uint32_t** hi, int* lo_bits) {
    int lb = k < 12 ? k : 12;
    *lo_bits = lb;
    uint64_t nlo = 1ull << lb, nhi = 1ull << (k - lb);
    std::lock_guard<std::mutex> tw_lock(c->sh->tw_mu);
    uint64_t klo = table_key(COSET_LO, curve, k, inverse), khi = table_key(COSET_HI, curve, k, inverse);
    auto it = c->twiddles.find(klo);
    if (it != c->twiddles.end()) {
        *lo = (const uint32_t*)it->second;
        *hi = (const uint32_t*)c->twiddles[khi];
        return;
    }
    void *pl = nullptr, *ph = nullptr;
    ZKM_CUDA(malloc_retry((void**)&pl, nlo * P::N * 4));
    c->twiddles[klo] = pl;
    ZKM_CUDA(malloc_retry((void**)&ph, nhi * P::N * 4));
    c->twiddles[khi] = ph;
    unsigned blocks = (unsigned)((nlo + nhi + 255) / 256);
    ZKM_LAUNCH(k_gen_coset<P>, blocks, 256, 0, s, (uint32_t*)pl, (uint32_t*)ph, nlo, nhi, inverse, k);
    ZKM_CUDA(cudaStreamSynchronize(s));
    *lo = (const uint32_t*)pl;
    *hi = (const uint32_t*)ph;
}

template <class P, int R>
static void launch_pass(Context* c, const NttPassArgs& a, cudaStream_t s) {
    typedef NttCfg<R, P::N> C;
    // Function attributes and occupancy are per device (one process may drive several GPUs) and this runs from
    // concurrent lanes: per-ordinal atomics, the CUDA calls themselves are idempotent.
    static std::atomic<int> per_sm_dev[64];
    const int ord = c->device & 63;
    int per_sm = per_sm_dev[ord].load(std::memory_order_acquire);
    if (per_sm == 0) {
        if (C::SMEM > 48 * 1024)
            ZKM_CUDA(cudaFuncSetAttribute(k_ntt_pass<P, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        ZKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ntt_pass<P, R>, C::NT, C::SMEM));
        if (per_sm < 1) per_sm = 1;
        per_sm_dev[ord].store(per_sm, std::memory_order_release);
    }
    uint64_t num_tiles = 1ull << (a.k - R);
    uint64_t ctas = (num_tiles + C::TILES - 1) / C::TILES;
    uint64_t cap = (uint64_t)c->sm_count * per_sm;
    unsigned grid = (unsigned)(ctas < cap ? ctas : cap);
    ZKM_LAUNCH((k_ntt_pass<P, R>), grid, C::NT, C::SMEM, s, a);
}

template <class P>
static void dispatch_pass(Context* c, int R, const NttPassArgs& a, cudaStream_t s) {
    switch (R) {
        case 3: launch_pass<P, 3>(c, a, s); break;
        case 4: launch_pass<P, 4>(c, a, s); break;
        case 5: launch_pass<P, 5>(c, a, s); break;
        case 6: launch_pass<P, 6>(c, a, s); break;
        case 7: launch_pass<P, 7>(c, a, s); break;
        case 8: launch_pass<P, 8>(c, a, s); break;
        case 9: launch_pass<P, 9>(c, a, s); break;
        case 10: launch_pass<P, 10>(c, a, s); break;
        case 11: launch_pass<P, 11>(c, a, s); break;
        case 12:
            if constexpr (NttMaxR<P>::value >= 12) { launch_pass<P, 12>(c, a, s); break; }
            [[fallthrough]];
        default: ZKM_FAIL(ZKM_ERR_ARG, "internal: bad NTT radix log %d", R);
    }
}

template <class P>
static void ntt_run_t(Context* c, int curve, const uint64_t* d_in, uint64_t* d_out, uint32_t log_n, int inverse,
                      int coset, cudaStream_t s) {
    const int k = (int)log_n;
    if (k > FrRoots<P>::TWO_ADICITY)
        ZKM_FAIL(ZKM_ERR_DOMAIN, "log_n %d exceeds the two-adicity %d of Fr", k, FrRoots<P>::TWO_ADICITY);
    if (k > 30) ZKM_FAIL(ZKM_ERR_ARG, "log_n %d: domains above 2^30 are not supported by this build", k);
    const uint32_t* in = (const uint32_t*)d_in;
    uint32_t* out = (uint32_t*)d_out;
    if (k <= 2) {
        ZKM_LAUNCH(k_ntt_tiny<P>, 1, 32, 0, s, in, out, k, inverse, coset);
        return;
    }
    int maxR = c->opt.ntt_max_radix_log;
    if (maxR < 6) maxR = 6;
    if (maxR > NttMaxR<P>::value) maxR = NttMaxR<P>::value;
    int passes = (k + maxR - 1) / maxR;
    int Rs[16];
    {   // balanced split, every pass >= 3 stages
        int base = k / passes, extra = k % passes;
        for (int i = 0; i < passes; i++) Rs[i] = base + (i < extra ? 1 : 0);
    }
    const uint64_t n = 1ull << k;
    // the last pass is out of place; when the caller wants in == out run it through scratch
    uint32_t* work = nullptr;       // buffer holding the in-place passes
    uint32_t* final_dst = out;
    bool copy_back = false;
    if (passes == 1) {
        if (in == out) {
            work = nullptr;
            final_dst = (uint32_t*)c->ntt_a.get(n * P::N * 4);
            copy_back = true;
        }
    } else {
        if (in == out) {
            work = out;  // in-place passes directly on the caller's buffer
            final_dst = (uint32_t*)c->ntt_a.get(n * P::N * 4);
            copy_back = true;
        } else {
            work = (uint32_t*)c->ntt_a.get(n * P::N * 4);  // first pass reads `in`, writes scratch
        }
    }
    const uint32_t *clo = nullptr, *chi = nullptr;
    int clo_bits = 0;
    if (coset) get_coset<P>(c, curve, k, inverse, s, &clo, &chi, &clo_bits);
    NttPassArgs a;
    memset(&a, 0, sizeof(a));
    if (inverse && !coset) {
        // size_inv lives in a cached 5-element device table of the domain constants
        std::lock_guard<std::mutex> tw_lock(c->sh->tw_mu);
        uint64_t key = table_key(DOMAIN, curve, k, 0);
        auto it = c->twiddles.find(key);
        void* p;
        if (it == c->twiddles.end()) {
            ZKM_CUDA(malloc_retry((void**)&p, 5 * P::N * 4));
            c->twiddles[key] = p;
            ZKM_LAUNCH(k_domain_constants<P>, 1, 32, 0, s, (uint32_t*)p, k);
            ZKM_CUDA(cudaStreamSynchronize(s));
        } else {
            p = it->second;
        }
        a.size_inv = (const uint32_t*)p + 2 * P::N;
    }
    int s0 = 0;
    for (int pi = 0; pi < passes; pi++) {
        const int R = Rs[pi];
        a.k = k;
        a.s0 = s0;
        a.first = pi == 0;
        a.last = pi == passes - 1;
        a.in = a.first ? in : work;
        a.out = a.last ? final_dst : work;
        a.tw_tile = get_twiddles<P>(c, curve, R, inverse, s);
        a.tw_four_tile = a.last ? nullptr : get_tw4<P>(c, curve, k - s0, R, inverse, s);
        a.tw_four = (a.last || a.tw_four_tile) ? nullptr : get_twiddles<P>(c, curve, k - s0, inverse, s);
        a.coset_lo = clo;
        a.coset_hi = chi;
        a.coset_lo_bits = clo_bits;
        a.scale_in = (a.first && coset && !inverse) ? 1 : 0;
        a.scale_out = a.last ? (inverse ? (coset ? 1 : 2) : 0) : 0;
        dispatch_pass<P>(c, R, a, s);
        s0 += R;
    }
    if (copy_back) ZKM_CUDA(cudaMemcpyAsync(out, final_dst, n * P::N * 4, cudaMemcpyDeviceToDevice, s));
}

// R1CStoQAP::witness_map on device-resident evaluation vectors a, b, c (each 2^k elements, consumed):
// h = coset_ifft( (coset_fft(ifft a) * coset_fft(ifft b) - coset_fft(ifft c)) / Z_H(g) ).
template <class P>
static void witness_map_t(Context* c, int curve, uint64_t* d_a, uint64_t* d_b, uint64_t* d_c, uint32_t log_n,
                          uint64_t* d_h, cudaStream_t s) {
    const uint64_t n = 1ull << log_n;
    uint64_t* tmp = (uint64_t*)c->ntt_b.get(n * P::N * 4);
    uint64_t* vecs[3] = {d_a, d_b, d_c};
    for (int v = 0; v < 3; v++) {
        ntt_run_t<P>(c, curve, vecs[v], tmp, log_n, 1, 0, s);   // ifft
        ntt_run_t<P>(c, curve, tmp, vecs[v], log_n, 0, 1, s);   // coset_fft
    }
    void* zinv;
    {
        std::lock_guard<std::mutex> tw_lock(c->sh->tw_mu);
        uint64_t key = table_key(VANISH, curve, (int)log_n, 0);
        auto it = c->twiddles.find(key);
        if (it == c->twiddles.end()) {
            ZKM_CUDA(malloc_retry((void**)&zinv, P::N * 4));
            c->twiddles[key] = zinv;
            ZKM_LAUNCH(k_vanishing_inv<P>, 1, 32, 0, s, (uint32_t*)zinv, (int)log_n);
            ZKM_CUDA(cudaStreamSynchronize(s));
        } else {
            zinv = it->second;
        }
    }
    uint64_t blocks = (n + 255) / 256;
    uint64_t cap = (uint64_t)c->sm_count * 8;
    ZKM_LAUNCH(k_qap_pointwise<P>, (unsigned)(blocks < cap ? blocks : cap), 256, 0, s, (uint32_t*)d_a, (const uint32_t*)d_b,
               (const uint32_t*)d_c, (const uint32_t*)zinv, n);
    ntt_run_t<P>(c, curve, d_a, d_h, log_n, 1, 1, s);           // coset_ifft
}

template <class P>
static void fr_into_repr_t(Context* c, const uint64_t* d_in, uint64_t* d_out, uint64_t n, cudaStream_t s) {
    if (n == 0) return;
    uint64_t blocks = (n + 255) / 256, cap = (uint64_t)c->sm_count * 8;
    ZKM_LAUNCH(k_fr_into_repr<P>, (unsigned)(blocks < cap ? blocks : cap), 256, 0, s, (const uint32_t*)d_in, (uint32_t*)d_out, n);
}

}  // namespace zkm

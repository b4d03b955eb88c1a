// zkm_msm_curve.cuh -- the curve-dependent MSM kernels (bucket accumulation K4, fold levels,
// window reduction K5, normalisation) as templates over the coordinate field; each
// zkm_msm_g{1,2}_{bls,bn}.cu instantiates them for one group.  See zkm_msm.cu for the pipeline and
// the upstream functions it replaces (ark-ec 0.3.0 src/msm/variable_base.rs,
// src/models/short_weierstrass_jacobian.rs).
#pragma once
#include "zkm_msm.cuh"
#include "zkm_msm_affine.cuh"
#include "zkm_msm_quad.cuh"

namespace zkm {

// ---------------------------------------------------------------------------------- curve tags
struct G1Bls { typedef Bls12_381_Fq F; static constexpr int SCALAR_BITS = 255; };
struct G2Bls { typedef Bls12_381_Fq2 F; static constexpr int SCALAR_BITS = 255; };
struct G1Bn { typedef Bn254_Fq F; static constexpr int SCALAR_BITS = 254; };
struct G2Bn { typedef Bn254_Fq2 F; static constexpr int SCALAR_BITS = 254; };
// BW6-761 (ark-bw6-761 0.3.0; /root/reference/benches/groth16.rs:24-29): G1 (y^2 = x^3 - 1) and G2 (y^2 = x^3 + 4)
// are both curves over the 761-bit Fq, scalars are 377 bits.  a = 0 and b never enters the group law, so the two
// groups share every kernel; only the synthetic-base generator differs.
struct G1Bw6 { typedef Bw6_761_Fq F; static constexpr int SCALAR_BITS = 377; };
struct G2Bw6 { typedef Bw6_761_Fq F; static constexpr int SCALAR_BITS = 377; };

template <class G> __device__ void load_generator(typename G::F& x, typename G::F& y);
template <> inline __device__ void load_generator<G1Bls>(Bls12_381_Fq& x, Bls12_381_Fq& y) {
    for (int i = 0; i < 12; i++) { x.l[i] = BLS12_381_G1_X[i]; y.l[i] = BLS12_381_G1_Y[i]; }
}
template <> inline __device__ void load_generator<G2Bls>(Bls12_381_Fq2& x, Bls12_381_Fq2& y) {
    for (int i = 0; i < 12; i++) {
        x.c0.l[i] = BLS12_381_G2_X0[i]; x.c1.l[i] = BLS12_381_G2_X1[i];
        y.c0.l[i] = BLS12_381_G2_Y0[i]; y.c1.l[i] = BLS12_381_G2_Y1[i];
    }
}
template <> inline __device__ void load_generator<G1Bn>(Bn254_Fq& x, Bn254_Fq& y) {
    for (int i = 0; i < 8; i++) { x.l[i] = BN254_G1_X[i]; y.l[i] = BN254_G1_Y[i]; }
}
template <> inline __device__ void load_generator<G2Bn>(Bn254_Fq2& x, Bn254_Fq2& y) {
    for (int i = 0; i < 8; i++) {
        x.c0.l[i] = BN254_G2_X0[i]; x.c1.l[i] = BN254_G2_X1[i];
        y.c0.l[i] = BN254_G2_Y0[i]; y.c1.l[i] = BN254_G2_Y1[i];
    }
}

template <> inline __device__ void load_generator<G1Bw6>(Bw6_761_Fq& x, Bw6_761_Fq& y) {
    for (int i = 0; i < 24; i++) { x.l[i] = BW6_761_G1_X[i]; y.l[i] = BW6_761_G1_Y[i]; }
}
template <> inline __device__ void load_generator<G2Bw6>(Bw6_761_Fq& x, Bw6_761_Fq& y) {
    for (int i = 0; i < 24; i++) { x.l[i] = BW6_761_G2_X[i]; y.l[i] = BW6_761_G2_Y[i]; }
}

template <class F>
__device__ __forceinline__ XYZZ<F> ld_xyzz(const XYZZ<F>* p) {
    XYZZ<F> r;
    const char* b = reinterpret_cast<const char*>(p);
    r.X = CoordIO<F>::ld_plain(b);
    r.Y = CoordIO<F>::ld_plain(b + CoordIO<F>::BYTES);
    r.ZZ = CoordIO<F>::ld_plain(b + 2 * CoordIO<F>::BYTES);
    r.ZZZ = CoordIO<F>::ld_plain(b + 3 * CoordIO<F>::BYTES);
    return r;
}
template <class F>
__device__ __forceinline__ void st_xyzz(XYZZ<F>* p, const XYZZ<F>& v) {
    char* b = reinterpret_cast<char*>(p);
    CoordIO<F>::st(b, v.X);
    CoordIO<F>::st(b + CoordIO<F>::BYTES, v.Y);
    CoordIO<F>::st(b + 2 * CoordIO<F>::BYTES, v.ZZ);
    CoordIO<F>::st(b + 3 * CoordIO<F>::BYTES, v.ZZZ);
}


// Out-of-line group law for the cold kernels (reduction tail, input generation): one copy of each
// formula per translation unit instead of one per call site keeps ptxas time and code size down.
template <class F> __device__ __noinline__ void xyzz_add_ni(XYZZ<F>& p, const XYZZ<F>& q) { xyzz_add(p, q); }
template <class F> __device__ __noinline__ void xyzz_dbl_ni(XYZZ<F>& p) { xyzz_dbl(p); }
template <class F> __device__ __noinline__ void xyzz_madd_ni(XYZZ<F>& p, const F& x, const F& y) { xyzz_madd(p, x, y); }
template <class F>
__device__ XYZZ<F> xyzz_mul_u64_ni(const F& x, const F& y, uint64_t k) {
    XYZZ<F> r = XYZZ<F>::identity();
    for (int b = 63; b >= 0; b--) {
        xyzz_dbl_ni(r);
        if ((k >> b) & 1) xyzz_madd_ni(r, x, y);
    }
    return r;
}
template <class F>
__device__ __noinline__ bool xyzz_to_affine_ni(const XYZZ<F>& p, F& x, F& y) { return xyzz_to_affine(p, x, y); }

// ---------------------------------------------------------------------------------- K4 accumulation
// CTAs per SM the accumulation kernel is compiled for: two (up to 255 registers).  Four (128 registers, 376 instead of 288
// bytes of stack for the 381-bit field) was measured in round 2 and is SLOWER: accumulate 3.19 -> 3.64 ms at 2^21, 6.74 -> 7.22
// at 2^20, 4.8 -> 5.0 at 2^24, no gain for concurrent small MSMs -- the extra spills sit on the dependent chain.
template <class F> struct AccumMinBlocks { static constexpr int value = 2; };
template <class F>
__global__ void __launch_bounds__(128, AccumMinBlocks<F>::value)
k_accum_affine(const char* __restrict__ bases, const uint32_t* __restrict__ idx, const TaskList tl, XYZZ<F>* __restrict__ out) {
    constexpr int CB = CoordIO<F>::BYTES;
    const uint32_t* __restrict__ tstart = tl.tstart;
    const uint32_t* __restrict__ tlen = tl.tlen;
    const uint32_t* __restrict__ order = tl.order;
    const uint32_t T = tl.tbase[tl.K];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
        const uint32_t task = order[t];
        const uint32_t s = tstart[task], len = tlen[task];
        // idx == nullptr: the lists are affine arrays written by the pairwise levels (position = entry,
        // no sign, identity markers possible)
        uint32_t id = idx ? idx[s] : s;
        const char* p = bases + (size_t)(id & 0x7fffffffu) * (2 * CB);
        F nx = CoordIO<F>::ld_gather(p), ny = CoordIO<F>::ld_gather(p + CB);
        uint32_t nsign = id >> 31;
        XYZZ<F> acc = XYZZ<F>::identity();
        for (uint32_t j = 0; j < len; j++) {
            F cx = nx, cy = ny;
            uint32_t csign = nsign;
            if (j + 1 < len) {
                id = idx ? idx[s + j + 1] : s + j + 1;
                p = bases + (size_t)(id & 0x7fffffffu) * (2 * CB);
                nx = CoordIO<F>::ld_gather(p);
                ny = CoordIO<F>::ld_gather(p + CB);
                nsign = id >> 31;
            }
            F my = neg(cy);
            if (csign) cy = my;
            if (idx || !aff_is_identity(cx)) xyzz_madd(acc, cx, cy);
        }
        st_xyzz(out + task, acc);
    }
}

// ---------------------------------------------------------------------------------- K5 reduction
// The G1 kernels inline the group law at the few hot call sites of the reduction (registers instead of
// local memory); G2 keeps the out-of-line copies: inlining them was measured (2^16 b_g2 MSM 8.9 -> 8.0 ms)
// but costs 4 more minutes of ptxas time per G2 translation unit.
template <class F> struct InlineLaw { static constexpr bool value = sizeof(F) <= 48; };
template <class F>
__device__ __forceinline__ void add_sel(XYZZ<F>& p, const XYZZ<F>& q) {
    if (InlineLaw<F>::value) xyzz_add(p, q); else xyzz_add_ni(p, q);
}
template <class F>
__device__ __forceinline__ void dbl_sel(XYZZ<F>& p) {
    if (InlineLaw<F>::value) xyzz_dbl(p); else xyzz_dbl_ni(p);
}

// Thread (w, t) owns buckets b in [t*g, (t+1)*g) of window w:  sum_b (b+1) S_b = acc + lo * run with
// run = sum S_b, acc = sum (b - lo + 1) S_b (running sums from the top), lo = t*g.
// HIER: store `run` next to `acc` instead of adding lo * run here.  The offsets lo = t * g are then applied by running
// the SAME reduction over the run sums one level up (sum_t t * run_t, see OpsImpl::reduce): no per-thread scalar
// multiplication -- 19 doublings + ~10 additions per thread at c = 20, 17 % on top of the 2 g additions of the loop
// (55 % at the g = 16 of a 2^21-point shard).
template <class F, bool HIER>
__global__ void __launch_bounds__(128)   // (128, 3) was measured: the register cap spills, 15.8 -> 18.1 ms at 2^24
k_bucket_reduce(const XYZZ<F>* __restrict__ items, const uint32_t* __restrict__ off, const uint32_t* __restrict__ cnt,
                uint32_t W, uint32_t B, uint32_t g, XYZZ<F>* __restrict__ contrib, XYZZ<F>* __restrict__ run_out) {
    const uint32_t per_w = B / g;
    const uint32_t gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= W * per_w) return;
    const uint32_t w = gid / per_w, t = gid % per_w;
    const uint32_t lo = t * g;
    XYZZ<F> run, acc;
    if (InlineLaw<F>::value) {
        // ONE inlined addition site for both `run += S` and `acc += run`: two inlined copies of add-2008-s (14
        // unrolled products, ~80 KB each) do not fit the instruction cache and the loop then streams its code from
        // L2 (ncu: `no_instruction` was the second-largest stall).  The operands rotate through three register sets
        // U += V with W as the bystander:  A (run += S): U = run, V = S, W = acc;  B (acc += run): U = acc, V = run.
        XYZZ<F> U = XYZZ<F>::identity(), V = XYZZ<F>::identity(), Wt = XYZZ<F>::identity();
        bool didA = false;
        for (uint32_t step = 0; step < 2 * g; step++) {
            if ((step & 1u) == 0) {
                const uint32_t k = w * B + (lo + g - 1 - (step >> 1));
                if (!cnt[k]) continue;          // empty bucket: no A step, layout stays U = acc, V = run
                Wt = U;                         // acc steps aside
                U = V;                          // run becomes the destination
                V = ld_xyzz(items + off[k]);
                didA = true;
            } else if (didA) {
                V = U;                          // run (updated) becomes the source
                U = Wt;                         // acc the destination
                didA = false;
            }
            xyzz_add(U, V);
        }
        acc = U;
        run = V;
    } else {
        run = XYZZ<F>::identity();
        acc = XYZZ<F>::identity();
        for (uint32_t b = lo + g; b-- > lo;) {
            const uint32_t k = w * B + b;
            if (cnt[k]) {
                XYZZ<F> S = ld_xyzz(items + off[k]);
                xyzz_add_ni(run, S);
            }
            xyzz_add_ni(acc, run);
        }
    }
    if (HIER) {
        st_xyzz(run_out + gid, run);
    } else if (lo != 0 && !run.is_identity()) {
        XYZZ<F> r = XYZZ<F>::identity();
        for (int bit = 31 - __clz(lo); bit >= 0; bit--) {
            xyzz_dbl_ni(r);
            if ((lo >> bit) & 1) xyzz_add_ni(r, run);
        }
        xyzz_add_ni(acc, r);
    }
    st_xyzz(contrib + gid, acc);
}

// Upper levels of the hierarchical reduction: the same running sums over a DENSE array of XYZZ points (the run sums
// of the level below, bucket b = in[w * stride_w + first + b], b < Bp), four lanes per chain (little work, all
// latency).  acc_out[w * per2 + t] = sum_{b in chunk t} (b - t g + 1) S_b,  run_out[...] = sum_{b in chunk t} S_b.
template <class F>
__global__ void __launch_bounds__(128)
k_reduce_dense_quad(const XYZZ<F>* __restrict__ in, uint32_t stride_w, uint32_t first, uint32_t Bp, uint32_t g, uint32_t per2,
                    uint32_t W, XYZZ<F>* __restrict__ acc_out, XYZZ<F>* __restrict__ run_out) {
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    if (gid >= W * per2) return;
    const int q = threadIdx.x & 3;
    const uint32_t mask = 0xfu << (threadIdx.x & 28);
    const uint32_t w = gid / per2, t = gid % per2;
    const uint32_t lo = t * g, hi = lo + g < Bp ? lo + g : Bp;
    XYZZ<F> run = XYZZ<F>::identity(), acc = XYZZ<F>::identity();
    for (uint32_t b = hi; b-- > lo;) {
        XYZZ<F> S = ld_xyzz(in + (size_t)w * stride_w + first + b);
        xyzz_add_quad(run, S, q, mask);
        xyzz_add_quad(acc, run, q, mask);
    }
    if (q == 0) {
        st_xyzz(acc_out + gid, acc);
        st_xyzz(run_out + gid, run);
    }
}

// window total = A_0 + G_0 (A_1 + G_1 (A_2 + ...)) with A_k = sums[k * W + w] (the plain sums of the levels' acc
// arrays) and G_k = 2^lg[k] the chunk length of level k: log2(B) doublings per window in all.  One quad per window.
struct ReduceLevels {
    int n;
    int lg[16];
};
template <class F>
__global__ void __launch_bounds__(128) k_window_combine(const XYZZ<F>* __restrict__ sums, ReduceLevels lv, uint32_t W,
                                                        XYZZ<F>* __restrict__ wsum) {
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    if (w >= W) return;
    const int q = threadIdx.x & 3;
    const uint32_t mask = 0xfu << (threadIdx.x & 28);
    XYZZ<F> v = ld_xyzz(sums + (size_t)(lv.n - 1) * W + w);
    for (int k = lv.n - 2; k >= 0; k--) {
        for (int i = 0; i < lv.lg[k]; i++) xyzz_dbl_quad(v, q, mask);
        XYZZ<F> a = ld_xyzz(sums + (size_t)k * W + w);
        xyzz_add_quad(v, a, q, mask);
    }
    if (q == 0) st_xyzz(wsum + w, v);
}

// result record: x, y (Montgomery affine), then one u64 flag: 0 = finite point, 1 = point at infinity (x = 0, y = 1
// like ark-ec GroupAffine::zero()), 2 = INVALID: a scalar of the MSM was not canonical (bits at or above the modulus
// width) -- the host entry points turn it into ZKM_ERR_SCALAR_RANGE, the *_device ones document it (zkm_b200.h)
template <class F>
__device__ void write_result(uint64_t* out, const XYZZ<F>& p, bool invalid = false) {
    constexpr int CB = CoordIO<F>::BYTES;
    F x, y;
    bool ok = !invalid && xyzz_to_affine_ni(p, x, y);
    if (!ok) {
        x = F::zero();
        y = F::one();
    }
    CoordIO<F>::st(out, x);
    CoordIO<F>::st(reinterpret_cast<char*>(out) + CB, y);
    out[2 * CB / 8] = invalid ? 2ull : (ok ? 0ull : 1ull);
}

// ---- lane-cooperative group law for the serial tails (zkm_msm_quad.cuh) and the Horner combine
template <class F>
__global__ void k_msm_final(const XYZZ<F>* __restrict__ wsum, int W, int c, const uint32_t* __restrict__ flags,
                            uint64_t* __restrict__ out) {
    if (threadIdx.x >= 4 || blockIdx.x != 0) return;
    const int q = threadIdx.x;     // four cooperating lanes, identical state
    const uint32_t mask = 0xfu;
    XYZZ<F> total = XYZZ<F>::identity();
    for (int w = W - 1; w >= 0; w--) {
        if (w != W - 1)
            for (int k = 0; k < c; k++) xyzz_dbl_quad_inl(total, q, mask);   // the 240-doubling chain: keep it in registers
        XYZZ<F> s = ld_xyzz(wsum + w);
        xyzz_add_quad(total, s, q, mask);
    }
    if (q == 0) write_result<F>(out, total, flags && flags[1] != 0);
}

// Quad versions of the tail kernels, used when the MSM is small (few thousand chains, all latency): one
// chain per FOUR lanes.  Same arguments and results as the scalar kernels above.
template <class F>
__global__ void __launch_bounds__(128)
k_bucket_reduce_quad(const XYZZ<F>* __restrict__ items, const uint32_t* __restrict__ off, const uint32_t* __restrict__ cnt,
                     uint32_t W, uint32_t B, uint32_t g, XYZZ<F>* __restrict__ contrib) {
    const uint32_t per_w = B / g;
    const uint32_t gid = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    if (gid >= W * per_w) return;
    const int q = threadIdx.x & 3;
    const uint32_t mask = 0xfu << (threadIdx.x & 28);
    const uint32_t w = gid / per_w, t = gid % per_w;
    const uint32_t lo = t * g;
    XYZZ<F> run = XYZZ<F>::identity(), acc = XYZZ<F>::identity();
    for (uint32_t b = lo + g; b-- > lo;) {
        const uint32_t k = w * B + b;
        if (cnt[k]) {
            XYZZ<F> S = ld_xyzz(items + off[k]);
            xyzz_add_quad(run, S, q, mask);
        }
        xyzz_add_quad(acc, run, q, mask);
    }
    if (lo != 0 && !run.is_identity()) {
        XYZZ<F> r = XYZZ<F>::identity();
        for (int bit = 31 - __clz(lo); bit >= 0; bit--) {
            xyzz_dbl_quad(r, q, mask);
            if ((lo >> bit) & 1) xyzz_add_quad(r, run, q, mask);
        }
        xyzz_add_quad(acc, r, q, mask);
    }
    if (q == 0) st_xyzz(contrib + gid, acc);
}

// block = 128 threads = 32 quads
template <class F>
__global__ void __launch_bounds__(128) k_window_sum_quad(const XYZZ<F>* __restrict__ in, uint32_t per_w, uint32_t chunk,
                                                         uint32_t nslices, XYZZ<F>* __restrict__ out) {
    __shared__ XYZZ<F> sh[32];
    const uint32_t w = blockIdx.x / nslices, slice = blockIdx.x % nslices;
    const uint32_t begin = slice * chunk;
    const uint32_t end = begin + chunk < per_w ? begin + chunk : per_w;
    const int q = threadIdx.x & 3;
    const uint32_t quad = threadIdx.x >> 2;
    const uint32_t mask = 0xfu << (threadIdx.x & 28);
    XYZZ<F> acc = XYZZ<F>::identity();
    for (uint32_t t = begin + quad; t < end; t += 32) {
        XYZZ<F> v = ld_xyzz(in + (size_t)w * per_w + t);
        xyzz_add_quad(acc, v, q, mask);
    }
    if (q == 0) sh[quad] = acc;
    __syncthreads();
    for (uint32_t s = 16; s > 0; s >>= 1) {
        if (quad < s) {
            XYZZ<F> a = sh[quad], b = sh[quad + s];
            xyzz_add_quad(a, b, q, mask);
            if (q == 0) sh[quad] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) st_xyzz(out + blockIdx.x, sh[0]);
}

// Plain sums of the acc arrays of ALL levels of the hierarchical reduction (level 0: W x per_w records, the upper
// levels a quarter of the one below each) in two launches: stage 1, one CTA per slice of <= 256 records (32 quads: 8
// dependent additions + a 5-step tree); stage 2, one CTA per (level, window) over its <= 32 slice sums.  ~170 us at
// 2^24 / c = 20; round 2a summed level 0 with a two-step single-lane tree (0.50 ms) and every upper level with one
// 32-quad CTA per window (64 dependent additions on the largest level: 0.71 ms).
constexpr uint32_t ZKM_SUM_SLICE = 256;
struct SumJobs {
    int n;                  // levels
    uint32_t off[17];       // record offset of the level's acc array from `base`
    uint32_t per[17];       // records per window at that level
    uint32_t nsl[17];       // slices per window
    uint32_t cta0[18];      // first stage-1 CTA of the level (cta0[n] = total)
    uint32_t sl0[17];       // first slice-sum record of the level
};
// sh[0] = sum of the 32 quads' partial sums (one CTA of 128 threads).  Written out on the shared array itself and after
// every address has been computed: the same tree behind a helper taking the array as a pointer, with the destination
// index read from the parameter struct after the calls, returned wrong points on sm_100a / nvcc 12.9 (caught by
// tools/microbench/sums_check.cu against a serial sum; the bisection variants are in that file).
#define ZKM_CTA_TREE_SUM_QUAD(sh, acc, quad, q, mask)                    \
    do {                                                                 \
        if ((q) == 0) (sh)[quad] = (acc);                                \
        __syncthreads();                                                 \
        for (uint32_t s2_ = 16; s2_ > 0; s2_ >>= 1) {                    \
            if ((quad) < s2_) {                                          \
                XYZZ<F> a_ = (sh)[quad], b_ = (sh)[(quad) + s2_];        \
                xyzz_add_quad(a_, b_, q, mask);                          \
                if ((q) == 0) (sh)[quad] = a_;                           \
            }                                                            \
            __syncthreads();                                             \
        }                                                                \
    } while (0)
template <class F>
__global__ void __launch_bounds__(128) k_sums_stage1(const XYZZ<F>* __restrict__ base, SumJobs jb, uint32_t W,
                                                     XYZZ<F>* __restrict__ slice_sums) {
    __shared__ XYZZ<F> sh[32];
    int L = 0;
    while (L + 1 < jb.n && blockIdx.x >= jb.cta0[L + 1]) L++;
    const uint32_t r = blockIdx.x - jb.cta0[L];
    const uint32_t w = r / jb.nsl[L], sl = r % jb.nsl[L];
    const uint32_t per = jb.per[L];
    const XYZZ<F>* in = base + jb.off[L] + (size_t)w * per;
    XYZZ<F>* dst = slice_sums + jb.sl0[L] + (size_t)w * jb.nsl[L] + sl;
    const uint32_t lo = sl * ZKM_SUM_SLICE, hi = lo + ZKM_SUM_SLICE < per ? lo + ZKM_SUM_SLICE : per;
    const int q = threadIdx.x & 3;
    const uint32_t quad = threadIdx.x >> 2;
    const uint32_t mask = 0xfu << (threadIdx.x & 28);
    XYZZ<F> acc = XYZZ<F>::identity();
    for (uint32_t t = lo + quad; t < hi; t += 32) {
        XYZZ<F> v = ld_xyzz(in + t);
        xyzz_add_quad(acc, v, q, mask);
    }
    ZKM_CTA_TREE_SUM_QUAD(sh, acc, quad, q, mask);
    if (threadIdx.x == 0) st_xyzz(dst, sh[0]);
}
// block (level, w): sums[level * W + w] = sum of the level's slice sums of window w
template <class F>
__global__ void __launch_bounds__(128) k_sums_stage2(const XYZZ<F>* __restrict__ slice_sums, SumJobs jb, uint32_t W,
                                                     XYZZ<F>* __restrict__ sums) {
    __shared__ XYZZ<F> sh[32];
    const uint32_t w = blockIdx.x % W, L = blockIdx.x / W;
    const uint32_t nsl = jb.nsl[L];
    const XYZZ<F>* in = slice_sums + jb.sl0[L] + (size_t)w * nsl;
    XYZZ<F>* dst = sums + (size_t)L * W + w;
    const int q = threadIdx.x & 3;
    const uint32_t quad = threadIdx.x >> 2;
    const uint32_t mask = 0xfu << (threadIdx.x & 28);
    XYZZ<F> acc = XYZZ<F>::identity();
    for (uint32_t t = quad; t < nsl; t += 32) {
        XYZZ<F> v = ld_xyzz(in + t);
        xyzz_add_quad(acc, v, q, mask);
    }
    ZKM_CTA_TREE_SUM_QUAD(sh, acc, quad, q, mask);
    if (threadIdx.x == 0) st_xyzz(dst, sh[0]);
}

// ---- fold: a bucket whose list was cut into several tasks holds several partial sums, consecutive records
// items[tbase[k] .. tbase[k] + tpb[k]); afterwards the bucket's sum is items[tbase[k]] (in place).  Which buckets:
// the class lists k_tasks_count wrote (zkm_msm.cu; S = ZKM_FOLD_SEG) -- the launch sequence is fixed, nothing is read
// back by the host.  n_lists = flags + 4: [0] |A|, [1] |L|, [2] segments, [3] |M|.
//   k_fold_seg   every segment of S consecutive partial sums of the M and L buckets -> stage[] (one quad each, all
//                segments of all buckets in parallel over the whole GPU)
//   k_fold_quad  class A: one quad sums the <= S partial sums in place; class M: one quad sums the <= S segment sums
//   k_fold_cta   class L (the top window of an MSM, the "ones" bucket of a Groth16 witness: thousands of partial
//                sums): one CTA per bucket over its segment sums, 64 quads take strided subsets, then a tree through
//                shared memory.  1 843 partial sums: 7 + 4 + 6 dependent additions.
template <class F>
__global__ void __launch_bounds__(128)
k_fold_seg(const XYZZ<F>* __restrict__ items, const uint32_t* __restrict__ tbase, const uint32_t* __restrict__ tpb,
           const uint32_t* __restrict__ seg_first, const uint32_t* __restrict__ segtab, XYZZ<F>* __restrict__ stage,
           const uint32_t* __restrict__ n_lists) {
    const uint32_t nS = n_lists[2];
    const int q = threadIdx.x & 3;
    const uint32_t mask = 0xfu << (threadIdx.x & 28);
    for (uint32_t sgm = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; sgm < nS; sgm += (gridDim.x * blockDim.x) >> 2) {
        const uint32_t k = segtab[2 * sgm], j = segtab[2 * sgm + 1];
        const uint32_t base = tbase[k] + j * ZKM_FOLD_SEG;
        uint32_t cnt = tpb[k] - j * ZKM_FOLD_SEG;
        if (cnt > ZKM_FOLD_SEG) cnt = ZKM_FOLD_SEG;
        XYZZ<F> acc = ld_xyzz(items + base);
        for (uint32_t i = 1; i < cnt; i++) {
            XYZZ<F> v = ld_xyzz(items + base + i);
            xyzz_add_quad(acc, v, q, mask);
        }
        if (q == 0) st_xyzz(stage + seg_first[k] + j, acc);
    }
}
template <class F>
__global__ void __launch_bounds__(128)
k_fold_quad(XYZZ<F>* __restrict__ items, const uint32_t* __restrict__ tbase, const uint32_t* __restrict__ tpb,
            const uint32_t* __restrict__ list, const uint32_t* __restrict__ seg_first, const XYZZ<F>* __restrict__ stage,
            uint32_t K, const uint32_t* __restrict__ n_lists) {
    const uint32_t nA = n_lists[0], nM = n_lists[3];
    const int q = threadIdx.x & 3;
    const uint32_t mask = 0xfu << (threadIdx.x & 28);
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; i < nA + nM; i += (gridDim.x * blockDim.x) >> 2) {
        const bool mid = i >= nA;
        const uint32_t k = mid ? list[K + 1 + (i - nA)] : list[i];
        const uint32_t dst = tbase[k];
        const XYZZ<F>* src = mid ? stage + seg_first[k] : items + dst;
        const uint32_t cnt = mid ? (tpb[k] + ZKM_FOLD_SEG - 1) / ZKM_FOLD_SEG : tpb[k];
        XYZZ<F> acc = ld_xyzz(src);
        for (uint32_t j = 1; j < cnt; j++) {
            XYZZ<F> v = ld_xyzz(src + j);
            xyzz_add_quad(acc, v, q, mask);
        }
        if (q == 0) st_xyzz(items + dst, acc);
    }
}
constexpr int ZKM_FOLD_NT = 256;
template <class F>
__global__ void __launch_bounds__(ZKM_FOLD_NT)
k_fold_cta(XYZZ<F>* __restrict__ items, const uint32_t* __restrict__ tbase, const uint32_t* __restrict__ tpb,
           const uint32_t* __restrict__ list, const uint32_t* __restrict__ seg_first, const XYZZ<F>* __restrict__ stage,
           uint32_t K, const uint32_t* __restrict__ n_lists) {
    constexpr uint32_t NQ = ZKM_FOLD_NT / 4;
    __shared__ XYZZ<F> sh[NQ];
    const uint32_t nL = n_lists[1];
    const int q = threadIdx.x & 3;
    const uint32_t quad = threadIdx.x >> 2;
    const uint32_t mask = 0xfu << (threadIdx.x & 28);
    for (uint32_t b = blockIdx.x; b < nL; b += gridDim.x) {
        const uint32_t k = list[K - 1 - b];
        const uint32_t nseg = (tpb[k] + ZKM_FOLD_SEG - 1) / ZKM_FOLD_SEG;
        const XYZZ<F>* in = stage + seg_first[k];
        XYZZ<F> acc = XYZZ<F>::identity();
        for (uint32_t j = quad; j < nseg; j += NQ) {
            XYZZ<F> v = ld_xyzz(in + j);
            xyzz_add_quad(acc, v, q, mask);
        }
        if (q == 0) sh[quad] = acc;
        __syncthreads();
        for (uint32_t st = NQ / 2; st > 0; st >>= 1) {
            if (quad < st) {
                XYZZ<F> a = sh[quad], o = sh[quad + st];
                xyzz_add_quad(a, o, q, mask);
                if (q == 0) sh[quad] = a;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) st_xyzz(items + tbase[k], sh[0]);
        __syncthreads();
    }
}

template <class F>
__global__ void k_write_identity(uint64_t* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    write_result<F>(out, XYZZ<F>::identity());
}

// sum of m affine records (2 coords + flag word)
template <class F>
__global__ void k_points_sum(const uint64_t* __restrict__ pts, uint64_t m, uint64_t* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    constexpr int CB = CoordIO<F>::BYTES;
    constexpr int REC = 2 * CB / 8 + 1;
    XYZZ<F> total = XYZZ<F>::identity();
    bool invalid = false;
    for (uint64_t i = 0; i < m; i++) {
        const uint64_t* r = pts + i * REC;
        if (r[REC - 1] == 2) invalid = true;   // a shard saw a non-canonical scalar
        if (r[REC - 1]) continue;
        // records are 8-byte aligned only (odd word count): read limb by limb
        F x, y;
        const uint32_t* r32 = reinterpret_cast<const uint32_t*>(r);
        uint32_t* xw = reinterpret_cast<uint32_t*>(&x);
        uint32_t* yw = reinterpret_cast<uint32_t*>(&y);
        for (int j = 0; j < CB / 4; j++) {
            xw[j] = r32[j];
            yw[j] = r32[CB / 4 + j];
        }
        xyzz_madd_ni(total, x, y);
    }
    write_result<F>(out, total, invalid);
}

// synthetic bases with known discrete logs: P_i = (a0 + i d) G, normalised per point
template <class G>
__global__ void __launch_bounds__(128) k_gen_progression(uint64_t a0, uint64_t d, uint64_t n, char* __restrict__ out) {
    typedef typename G::F F;
    constexpr int CB = CoordIO<F>::BYTES;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F gx, gy;
    load_generator<G>(gx, gy);
    XYZZ<F> p = xyzz_mul_u64_ni(gx, gy, a0 + i * d);
    F x, y;
    if (!xyzz_to_affine_ni(p, x, y)) {
        x = F::zero();
        y = F::one();
    }
    CoordIO<F>::st(out + i * 2 * CB, x);
    CoordIO<F>::st(out + i * 2 * CB + CB, y);
}


// window multiples of registered bases: table[w * n + i] = affine(2^(c w) P_i).  One-time cost at
// registration (W inversions per base); afterwards every window of an MSM lands in ONE bucket set and
// the Horner doubling chain of the window combine disappears.
template <class F>
__global__ void __launch_bounds__(128) k_precompute(const char* __restrict__ bases, const uint8_t* __restrict__ inf,
                                                    uint64_t n, int c, int W, char* __restrict__ table) {
    constexpr int CB = CoordIO<F>::BYTES;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = CoordIO<F>::ld(bases + i * 2 * CB), y = CoordIO<F>::ld(bases + i * 2 * CB + CB);
    const bool is_inf = inf && inf[i];
    XYZZ<F> p = xyzz_from_affine(x, y);
    for (int w = 0; w < W; w++) {
        if (w > 0) {
            for (int k = 0; k < c; k++) xyzz_dbl_ni(p);
            if (is_inf || !xyzz_to_affine_ni(p, x, y)) {
                x = F::zero();
                y = F::zero();
            }
        }
        char* dst = table + ((uint64_t)w * n + i) * 2 * CB;
        CoordIO<F>::st(dst, x);
        CoordIO<F>::st(dst + CB, y);
    }
}

// ---------------------------------------------------------------------------------- ops table
// Build-time split: for the 761-bit field the batched-affine pair kernels are compiled in their own translation unit
// (zkm_msm_bw6_pair.cu) so that ptxas works on the two halves of the group in parallel.  `external` = true makes
// OpsImpl<G>::make() take the pair launchers from there instead of instantiating them in the including unit.
template <class G> struct PairOpsProvider { static constexpr bool external = false; };
template <> struct PairOpsProvider<G1Bw6> { static constexpr bool external = true; };
template <> struct PairOpsProvider<G2Bw6> { static constexpr bool external = true; };
void bw6_pair_fwd(unsigned sm_count, uint64_t nT_bound, cudaStream_t s, int level0, const void* src, const uint32_t* idx,
                  const uint32_t* map, const uint32_t* off_out, uint32_t K, uint32_t m, void* pre, void* T, const void* xarr);
void bw6_pair_inv(unsigned sm_count, uint64_t nU_bound, cudaStream_t s, const uint32_t* off_out, uint32_t K, uint32_t m,
                  uint32_t m2, void* T, void* pre2);
void bw6_pair_bwd(unsigned sm_count, uint64_t nT_bound, cudaStream_t s, int level0, const void* src, const uint32_t* idx,
                  const uint32_t* map, const uint32_t* off_out, uint32_t K, uint32_t m, const void* pre, const void* Tinv,
                  void* dst);
void bw6_build_xarr(unsigned sm_count, cudaStream_t s, const void* bases, uint64_t n, void* xarr);

template <class G>
struct OpsImpl {
    typedef typename G::F F;
    static void accum_affine(unsigned grid, cudaStream_t s, const void* bases, const uint32_t* idx, TaskList tl, void* out) {
        ZKM_LAUNCH(k_accum_affine<F>, grid, 128, 0, s, (const char*)bases, idx, tl, (XYZZ<F>*)out);
    }
    static void fold(unsigned sm_count, cudaStream_t s, void* items, const uint32_t* tbase, const uint32_t* tpb,
                     const uint32_t* fold_list, const uint32_t* seg_first, const uint32_t* segtab, void* stage, uint32_t K,
                     uint32_t max_segs, const uint32_t* n_lists) {
        // persistent grids over device-side counts: CTAs without work exit at once
        auto quad_grid = [&](uint64_t quads) {
            const uint64_t need = (quads * 4 + 127) / 128, cap = (uint64_t)sm_count * 4;
            return (unsigned)(need < cap ? (need ? need : 1) : cap);
        };
        ZKM_LAUNCH(k_fold_seg<F>, quad_grid(max_segs), 128, 0, s, (const XYZZ<F>*)items, tbase, tpb, seg_first, segtab,
                   (XYZZ<F>*)stage, n_lists);
        ZKM_LAUNCH(k_fold_quad<F>, quad_grid(K), 128, 0, s, (XYZZ<F>*)items, tbase, tpb, fold_list, seg_first,
                   (const XYZZ<F>*)stage, K, n_lists);
        const unsigned gc = K < sm_count ? K : sm_count;
        ZKM_LAUNCH(k_fold_cta<F>, gc, ZKM_FOLD_NT, 0, s, (XYZZ<F>*)items, tbase, tpb, fold_list, seg_first, (const XYZZ<F>*)stage, K,
                   n_lists);
    }
    // persistent grids: exactly the co-resident CTAs (or fewer when the level is small)
    template <class K>
    static unsigned pair_grid(K kernel, int block, unsigned sm_count, uint64_t threads_needed, std::atomic<int>* cache) {
        int occ = cache->load(std::memory_order_acquire);     // written from concurrent lanes: same value, atomically
        if (occ == 0) {
            ZKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, 0));
            if (occ < 1) occ = 1;
            cache->store(occ, std::memory_order_release);
        }
        uint64_t need = (threads_needed + block - 1) / block;
        uint64_t cap = (uint64_t)sm_count * (unsigned)occ;
        return (unsigned)(need < cap ? (need ? need : 1) : cap);
    }
    static void pair_fwd(unsigned sm_count, uint64_t nT_bound, cudaStream_t s, int level0, const void* src, const uint32_t* idx,
                         const uint32_t* map, const uint32_t* off_out, uint32_t K, uint32_t m, void* pre, void* T, const void* xarr) {
        static std::atomic<int> occ0{0}, occ1{0};
        if (level0) {
            unsigned grid = pair_grid(k_pair_fwd<F, true>, 256, sm_count, nT_bound, &occ0);
            ZKM_LAUNCH((k_pair_fwd<F, true>), grid, 256, 0, s, (const char*)src, idx, map, off_out, K, m, (char*)pre, (char*)T, (const char*)xarr);
        } else {
            unsigned grid = pair_grid(k_pair_fwd<F, false>, 256, sm_count, nT_bound, &occ1);
            ZKM_LAUNCH((k_pair_fwd<F, false>), grid, 256, 0, s, (const char*)src, idx, map, off_out, K, m, (char*)pre, (char*)T, (const char*)xarr);
        }
    }
    // x coordinates of `n` bases in 64-byte slots (level-0 forward gathers); a no-op for fields without such a layout
    static void build_xarr(unsigned sm_count, cudaStream_t s, const void* bases, uint64_t n, void* xarr) {
        if (XArr<F>::SLOT == 0 || n == 0) return;
        uint64_t blocks = (n + 255) / 256, cap = (uint64_t)sm_count * 8;
        ZKM_LAUNCH(k_build_xarr<F>, (unsigned)(blocks < cap ? blocks : cap), 256, 0, s, (const char*)bases, n, (char*)xarr);
    }
    static void pair_inv(unsigned sm_count, uint64_t nU_bound, cudaStream_t s, const uint32_t* off_out, uint32_t K, uint32_t m,
                         uint32_t m2, void* T, void* pre2) {
        static std::atomic<int> occ{0};
        unsigned grid = pair_grid(k_inv_batch<F>, 128, sm_count, nU_bound, &occ);
        ZKM_LAUNCH(k_inv_batch<F>, grid, 128, 0, s, off_out, K, m, m2, (char*)T, (char*)pre2);
    }
    static void pair_bwd(unsigned sm_count, uint64_t nT_bound, cudaStream_t s, int level0, const void* src, const uint32_t* idx,
                         const uint32_t* map, const uint32_t* off_out, uint32_t K, uint32_t m, const void* pre, const void* Tinv,
                         void* dst) {
        static std::atomic<int> occ0{0}, occ1{0};
        if (level0) {
            unsigned grid = pair_grid(k_pair_bwd<F, true>, 128, sm_count, nT_bound, &occ0);
            ZKM_LAUNCH((k_pair_bwd<F, true>), grid, 128, 0, s, (const char*)src, idx, map, off_out, K, m, (const char*)pre, (const char*)Tinv, (char*)dst);
        } else {
            unsigned grid = pair_grid(k_pair_bwd<F, false>, 128, sm_count, nT_bound, &occ1);
            ZKM_LAUNCH((k_pair_bwd<F, false>), grid, 128, 0, s, (const char*)src, idx, map, off_out, K, m, (const char*)pre, (const char*)Tinv, (char*)dst);
        }
    }
    static void reduce(cudaStream_t s, const void* items, const uint32_t* off, const uint32_t* cnt, MsmPlan pl,
                       void* contrib, void* wsum, const uint32_t* flags, uint64_t* d_out) {
        const uint32_t RW = (uint32_t)pl.RW;
        uint32_t g = msm_reduce_group(pl.B, RW);
        uint32_t per_w = pl.B / g;
        // small reductions are chains of dependent additions: run them four lanes per chain
        const bool quad = (uint64_t)RW * per_w <= 16384;
        XYZZ<F>* stage = (XYZZ<F>*)contrib + (size_t)RW * per_w;
        if (quad) {
            unsigned rblocks = (RW * per_w * 4 + 127) / 128;
            ZKM_LAUNCH(k_bucket_reduce_quad<F>, rblocks, 128, 0, s, (const XYZZ<F>*)items, off, cnt, RW, pl.B, g,
                       (XYZZ<F>*)contrib);
            uint32_t nslices = (per_w + 255) / 256;
            if (nslices > 1) {
                uint32_t chunk = (per_w + nslices - 1) / nslices;
                ZKM_LAUNCH(k_window_sum_quad<F>, RW * nslices, 128, 0, s, (const XYZZ<F>*)contrib, per_w, chunk, nslices, stage);
                ZKM_LAUNCH(k_window_sum_quad<F>, RW, 128, 0, s, (const XYZZ<F>*)stage, nslices, nslices, 1u, (XYZZ<F>*)wsum);
            } else {
                ZKM_LAUNCH(k_window_sum_quad<F>, RW, 128, 0, s, (const XYZZ<F>*)contrib, per_w, per_w, 1u, (XYZZ<F>*)wsum);
            }
        } else {
            // Hierarchical running sums.  Level 0: one thread per chunk of g buckets -> acc (weights relative to the
            // chunk) and run (plain chunk sum).  The chunk offsets t * g are worth g * sum_t t * run_t: the same
            // reduction over the run sums (a dense array now, four lanes per chain), and so on until one chunk is left.
            XYZZ<F>* acc0 = (XYZZ<F>*)contrib;                       // RW * per_w
            XYZZ<F>* run0 = acc0 + (size_t)RW * per_w;
            XYZZ<F>* upper = run0 + (size_t)RW * per_w;              // acc / run of the upper levels, sums, slice sums
            unsigned rblocks = (RW * per_w + 127) / 128;
            ZKM_LAUNCH((k_bucket_reduce<F, true>), rblocks, 128, 0, s, (const XYZZ<F>*)items, off, cnt, RW, pl.B, g, acc0, run0);
            ReduceLevels lv;
            lv.n = 1;
            lv.lg[0] = 31 - __builtin_clz(g);
            size_t used = 0;                                           // records of `upper` in use
            auto take = [&](size_t n) { XYZZ<F>* p = upper + used; used += n; return p; };
            XYZZ<F>* sums = take((size_t)17 * RW);
            SumJobs jb;
            jb.n = 0;
            uint32_t n_cta = 0, n_sl = 0;
            auto add_level = [&](const XYZZ<F>* acc, uint32_t per) {
                const int L = jb.n++;
                jb.off[L] = (uint32_t)(acc - acc0);
                jb.per[L] = per;
                jb.nsl[L] = (per + ZKM_SUM_SLICE - 1) / ZKM_SUM_SLICE;
                jb.cta0[L] = n_cta;
                jb.sl0[L] = n_sl;
                n_cta += RW * jb.nsl[L];
                n_sl += RW * jb.nsl[L];
                jb.cta0[L + 1] = n_cta;
            };
            add_level(acc0, per_w);
            const XYZZ<F>* cur_run = run0;
            uint32_t cur_per = per_w;
            while (cur_per > 1) {
                const uint32_t Bp = cur_per - 1;                      // chunk t >= 1 has weight t: bucket b = t - 1
                // upper levels are pure latency (2 gk dependent additions each, little parallel work): short chunks.
                // Measured at 2^21 / c = 16 with gk = 8 and a last chunk of up to 32: 0.95 ms in these levels.
                uint32_t gk = 4;
                if (Bp <= 4) { gk = 1; while (gk < Bp) gk <<= 1; }    // last level: one chunk
                const uint32_t per2 = (Bp + gk - 1) / gk;
                XYZZ<F>* acck = take((size_t)RW * per2);
                XYZZ<F>* runk = take((size_t)RW * per2);
                ZKM_LAUNCH(k_reduce_dense_quad<F>, (RW * per2 * 4 + 127) / 128, 128, 0, s, cur_run, cur_per, 1u, Bp, gk, per2, RW,
                           acck, runk);
                add_level(acck, per2);
                lv.lg[lv.n] = 31 - __builtin_clz(gk);
                lv.n++;
                cur_run = runk;
                cur_per = per2;
            }
            XYZZ<F>* slice_sums = take(n_sl);
            ZKM_LAUNCH(k_sums_stage1<F>, n_cta, 128, 0, s, (const XYZZ<F>*)acc0, jb, RW, slice_sums);
            ZKM_LAUNCH(k_sums_stage2<F>, RW * (unsigned)jb.n, 128, 0, s, (const XYZZ<F>*)slice_sums, jb, RW, sums);
            ZKM_LAUNCH(k_window_combine<F>, (RW * 4 + 127) / 128, 128, 0, s, (const XYZZ<F>*)sums, lv, RW, (XYZZ<F>*)wsum);
        }
        ZKM_LAUNCH(k_msm_final<F>, 1, 32, 0, s, (const XYZZ<F>*)wsum, pl.RW, pl.c, flags, d_out);
    }
    static void write_identity(cudaStream_t s, uint64_t* d_out) { ZKM_LAUNCH(k_write_identity<F>, 1, 32, 0, s, d_out); }
    static void points_sum(cudaStream_t s, const uint64_t* pts, uint64_t m, uint64_t* d_out) {
        ZKM_LAUNCH(k_points_sum<F>, 1, 32, 0, s, pts, m, d_out);
    }
    static void gen_progression(cudaStream_t s, uint64_t a0, uint64_t d, uint64_t n, void* d_out) {
        if (n == 0) return;
        unsigned blocks = (unsigned)((n + 127) / 128);
        ZKM_LAUNCH(k_gen_progression<G>, blocks, 128, 0, s, a0, d, n, (char*)d_out);
    }
    static void precompute(cudaStream_t s, const void* bases, const uint8_t* inf, uint64_t n, int c, int W, void* table) {
        if (n == 0) return;
        unsigned blocks = (unsigned)((n + 127) / 128);
        ZKM_LAUNCH(k_precompute<F>, blocks, 128, 0, s, (const char*)bases, inf, n, c, W, (char*)table);
    }
    static CurveOps make(int curve, int group) {
        CurveOps o;
        o.curve = curve;
        o.group = group;
        o.scalar_bits = G::SCALAR_BITS;
        o.xyzz_bytes = sizeof(XYZZ<F>);
        o.accum_affine = accum_affine;
        o.fold = fold;
        o.accum_ctas_per_sm = AccumMinBlocks<F>::value;
        o.reduce = reduce;
        o.write_identity = write_identity;
        o.points_sum = points_sum;
        o.gen_progression = gen_progression;
        o.precompute = precompute;
        if constexpr (PairOpsProvider<G>::external) {
            o.pair_fwd = bw6_pair_fwd;
            o.build_xarr = bw6_build_xarr;
            o.pair_inv = bw6_pair_inv;
            o.pair_bwd = bw6_pair_bwd;
        } else {
            o.pair_fwd = pair_fwd;
            o.build_xarr = build_xarr;
            o.pair_inv = pair_inv;
            o.pair_bwd = pair_bwd;
        }
        o.xarr_slot = XArr<F>::SLOT;
        o.coord_bytes = CoordIO<F>::BYTES;
        return o;
    }
};

}  // namespace zkm

// zkm_msm.cu -- variable-base multi-scalar multiplication (G1 and G2) on sm_100a.
//
// Replaces ark-ec 0.3.0 VariableBaseMSM::multi_scalar_mul (src/msm/variable_base.rs) plus the
// final into_affine() (src/models/short_weierstrass_jacobian.rs); pin
// /root/reference/Cargo.lock:179-180; reached from /root/reference/benches/groth16.rs:115
// (ark-groth16 create_proof: h/l/a/b_g1 queries on G1, b_g2 query on G2) and from
// /root/reference/benches/marlin.rs:202,311 (KZG10::commit).  The result is the unique normalised
// affine point, so it is byte-identical to upstream's whatever the bucket schedule.
//
// Pipeline (a from-scratch design, not upstream's one-rayon-task-per-window loop):
//   K2  k_msm_digits<0>   signed c-bit digits of every scalar, histogram of (window, |digit|) buckets
//       scan              exclusive prefix sum of bucket sizes -> bucket offsets
//   K3  k_msm_digits<1>   scatter (point index | sign) into digit-sorted bucket lists
//       k_tasks_*         cut every bucket list into tasks of <= L entries, order tasks by length
//   K4  k_accum_affine    one thread per task: XYZZ accumulator += affine base (madd-2008-s), bases
//                         gathered with 128-bit loads; then k_fold_quad / k_fold_cta sum the partial
//                         sums of every bucket that was cut into several tasks (keeps skewed inputs --
//                         the 0/1-heavy Groth16 witness -- balanced without a special case and without
//                         a device read-back)
//   K5  k_bucket_reduce   running-sum reduction of each window in parallel slices, k_window_sum,
//       k_msm_final       Horner combine over windows (c doublings each) and normalisation to affine
// The integer pipe (Montgomery products) bounds K4; everything else is a few percent.  See DESIGN.md.
#include <math.h>

#include <cub/device/device_scan.cuh>

#include "zkm_msm.cuh"

namespace zkm {

// ---------------------------------------------------------------------------------- K2 / K3 digits

// MODE 0: histogram into counts[K] (all windows).  MODE 1: scatter (index | sign << 31) at cursor[key]++
// for window w_only: one launch per window keeps the write set (n x 4 B) inside the 126 MB L2, so the
// random 4-byte stores merge into full sectors before they reach HBM.
// SL = 32-bit limbs per scalar: 8 (BigInteger256) or 12 (BigInteger384, BW6-761).
// Bucket atomics.  Same-address atomics serialise: window 0 of a Groth16 witness is 45 % ones (7.5 M hits on one counter
// at 2^24: sort 4.1 -> 8.2 ms), the top window has few real bits, a registration with window multiples has ONE small
// bucket set for all windows.  On those windows the lanes that share the bucket of the warp's first entry are counted
// with one atomic (two ballots and a shuffle).  A full __match_any_sync on every window was measured and dropped: it
// doubled the cost of each window it ran on (sort 5.3 -> 7.0 ms at 2^24 uniform).
// `dig` (MODE 0, large MSMs): the histogram pass also writes every digit as one word, dig[w * n + i] = |d| | sign << 31
// (0: no entry), so that the per-window scatter launches (k_msm_scatter_row) read 4 bytes per scalar instead of
// decoding the whole 32-byte scalar again in each of the W launches (13 x 512 MB at 2^24 -> 0.9 GB written + read).
template <int MODE, int SL>
__global__ void __launch_bounds__(256) k_msm_digits(const uint32_t* __restrict__ scalars, const uint8_t* __restrict__ inf,
                                                    uint64_t n, MsmPlan pl, int w_only,
                                                    uint32_t* __restrict__ counts_or_cursor,
                                                    uint32_t* __restrict__ idx_out, uint32_t* __restrict__ flags,
                                                    uint32_t* __restrict__ dig, uint64_t i_begin, uint64_t i_end) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const bool write_dig = MODE == 0 && dig != nullptr;
    // warp-uniform loop over the scalars [i_begin, i_end) (the whole array, or one chunk of a ScalarFeed): the lanes of a
    // warp walk the digit loop together (lanes without a scalar carry zeros)
    for (uint64_t i0 = i_begin + (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < i_end; i0 += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = i0 + lane;
        const bool in_range = i < i_end;
        bool live = in_range && !(inf && inf[i]);
        uint32_t s[SL];
        uint32_t any = 0;
#pragma unroll
        for (int v = 0; v < SL; v++) s[v] = 0;
        if (live) {
            const uint4* sp = reinterpret_cast<const uint4*>(scalars + i * SL);
#pragma unroll
            for (int v = 0; v < SL / 4; v++) {
                uint4 a = __ldg(sp + v);
                s[4 * v] = a.x; s[4 * v + 1] = a.y; s[4 * v + 2] = a.z; s[4 * v + 3] = a.w;
                any |= a.x | a.y | a.z | a.w;
            }
        }
        live = live && any != 0;
        if (!write_dig && __ballot_sync(0xffffffffu, live) == 0) continue;   // (with `dig` the zero rows must still be written)
        const uint32_t mask = (1u << pl.c) - 1u;
        uint64_t buf = 0;
        int nb = 0, w = 0;
        uint32_t carry = 0;
#pragma unroll
        for (int limb = 0; limb <= SL; limb++) {
            if (limb < SL) {
                buf |= (uint64_t)s[limb] << nb;
                nb += 32;
            } else {
                nb = 64;  // flush: the rest of the buffer is zero extension
            }
            while (nb >= pl.c && w < pl.W && (MODE == 0 || w_only < 0 || w <= w_only)) {
                uint32_t d = ((uint32_t)buf & mask) + carry;
                buf >>= pl.c;
                nb -= pl.c;
                uint32_t sign = 0;
                carry = 0;
                if (d > pl.B) {
                    d = (1u << pl.c) - d;
                    sign = 1;
                    carry = 1;
                }
                if (write_dig && in_range) dig[(size_t)w * n + i] = d ? (d | (sign << 31)) : 0u;
                if (MODE == 0 || w_only < 0 || w == w_only) {      // uniform over the warp
                    const bool hit = d != 0;                        // (lanes without a scalar have d == 0)
                    const uint32_t key = (uint32_t)w * pl.key_stride + (d - 1);
                    const uint32_t entry = ((uint32_t)w * pl.idx_stride + pl.idx_base + (uint32_t)i) | (sign << 31);
                    if (w == 0 || w == pl.W - 1 || pl.key_stride == 0) {
                        const uint32_t hits = __ballot_sync(0xffffffffu, hit);
                        if (hits) {
                            const int l0 = __ffs(hits) - 1;
                            const uint32_t k0 = __shfl_sync(0xffffffffu, key, l0);
                            const uint32_t same = __ballot_sync(0xffffffffu, hit && key == k0);
                            uint32_t base = 0;
                            if ((int)lane == l0) base = atomicAdd(&counts_or_cursor[k0], (uint32_t)__popc(same));
                            if (MODE == 1) base = __shfl_sync(0xffffffffu, base, l0);
                            if (hit) {
                                if (key == k0) {
                                    if (MODE == 1) idx_out[base + (uint32_t)__popc(same & lt_mask)] = entry;
                                } else if (MODE == 0) {
                                    atomicAdd(&counts_or_cursor[key], 1u);
                                } else {
                                    idx_out[atomicAdd(&counts_or_cursor[key], 1u)] = entry;
                                }
                            }
                        }
                    } else if (hit) {
                        if (MODE == 0) atomicAdd(&counts_or_cursor[key], 1u);
                        else idx_out[atomicAdd(&counts_or_cursor[key], 1u)] = entry;
                    }
                }
                w++;
            }
        }
        // a canonical scalar is < r < 2^scalar_bits: any bit at or above that position (or a digit carry out of the
        // top window) marks the input as non-canonical -> flag word 2 of the result record / ZKM_ERR_SCALAR_RANGE
        if (MODE == 0 && live && (buf != 0 || carry != 0 || (s[SL - 1] >> (pl.scalar_bits & 31)) != 0)) atomicOr(&flags[1], 1u);
    }
}

// scatter of one window from its row of digit words (see k_msm_digits): idx_out[cursor[key]++] = entry.
// The lanes that share the bucket of the warp's first entry take their slots with ONE atomic (two ballots and a shuffle,
// not a __match_any_sync): skewed rows -- window 0 of a Groth16 witness is 45 % ones, the top window has few real bits --
// otherwise serialise millions of returning atomics on one counter (2^24 witness scalars: sort 4.1 -> 11.2 ms without it),
// and uniform rows pay almost nothing.
__global__ void __launch_bounds__(256) k_msm_scatter_row(const uint32_t* __restrict__ dig_row, uint64_t n, MsmPlan pl, int w,
                                                         uint32_t* __restrict__ cursor, uint32_t* __restrict__ idx_out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); i0 < n; i0 += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = i0 + lane;
        const uint32_t dw = i < n ? __ldg(dig_row + i) : 0u;
        const bool hit = dw != 0;
        const uint32_t hits = __ballot_sync(0xffffffffu, hit);
        if (hits == 0) continue;
        const uint32_t key = (uint32_t)w * pl.key_stride + ((dw & 0x7fffffffu) - 1u);
        const int l0 = __ffs(hits) - 1;
        const uint32_t k0 = __shfl_sync(0xffffffffu, key, l0);
        const uint32_t same = __ballot_sync(0xffffffffu, hit && key == k0);
        uint32_t base = 0;
        if ((int)lane == l0) base = atomicAdd(&cursor[k0], (uint32_t)__popc(same));
        base = __shfl_sync(0xffffffffu, base, l0);
        if (hit) {
            const uint32_t pos = key == k0 ? base + (uint32_t)__popc(same & lt_mask) : atomicAdd(&cursor[key], 1u);
            idx_out[pos] = ((uint32_t)w * pl.idx_stride + pl.idx_base + (uint32_t)i) | (dw & 0x80000000u);
        }
    }
}

// len_out[k] = ceil(len_in[k] / 2) with len_in[k] = off_in[k+1] - off_in[k]  (one pairwise level)
__global__ void k_pair_lens(const uint32_t* __restrict__ off_in, uint32_t K, uint32_t* __restrict__ len_out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < K) len_out[k] = (off_in[k + 1] - off_in[k] + 1) >> 1;
    else if (k == K) len_out[k] = 0;
}

// Where every output of a pairwise level finds its inputs: map[e] = first input position | (two inputs ? 1 << 31 : 0)
// for output e of bucket k, e in [off_out[k], off_out[k+1]): inputs off_in[k] + 2 (e - off_out[k]) and the next one (the
// odd element at the end of a list is passed through).  One thread walks ZKM_MAP_M consecutive outputs with a bucket
// cursor (one binary search per thread); the CTA's 8192 words go through shared memory so that the stores are
// contiguous (written straight from the threads -- 32 words 128 bytes apart per store -- the four launches of a
// 2^24-point MSM took 2.0 ms).  The pair kernels (zkm_msm_affine.cuh) then need neither offsets nor a search and can
// take the outputs in any order.
constexpr uint32_t ZKM_MAP_M = 32;
__global__ void __launch_bounds__(256) k_pair_map(const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out,
                                                  uint32_t K, uint32_t* __restrict__ map) {
    __shared__ uint32_t sh[256 * (ZKM_MAP_M + 1)];
    const uint32_t E = off_out[K];
    const uint32_t per_cta = 256 * ZKM_MAP_M;
    for (uint64_t base = (uint64_t)blockIdx.x * per_cta; base < E; base += (uint64_t)gridDim.x * per_cta) {
        const uint64_t e0_64 = base + (uint64_t)threadIdx.x * ZKM_MAP_M;
        if (e0_64 < E) {
            const uint32_t e0 = (uint32_t)e0_64, e1 = (E - e0 > ZKM_MAP_M) ? e0 + ZKM_MAP_M : E;
            uint32_t a = 0, b = K;  // off_out[a] <= e0 < off_out[b]
            while (b - a > 1) {
                const uint32_t mid = (a + b) >> 1;
                if (off_out[mid] <= e0) a = mid; else b = mid;
            }
            uint32_t k = a, lo = off_out[a], hi = off_out[a + 1];
            uint32_t ib = off_in[k], ie = off_in[k + 1];
            for (uint32_t e = e0; e < e1; e++) {
                while (e >= hi) {
                    k++;
                    lo = hi;
                    hi = off_out[k + 1];
                    ib = ie;
                    ie = off_in[k + 1];
                }
                const uint32_t i0 = ib + 2 * (e - lo);
                sh[threadIdx.x * (ZKM_MAP_M + 1) + (e - e0)] = i0 | ((i0 + 1 < ie) ? 0x80000000u : 0u);
            }
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < per_cta; i += 256) {
            const uint64_t e = base + i;
            if (e < E) map[e] = sh[(i / ZKM_MAP_M) * (ZKM_MAP_M + 1) + (i % ZKM_MAP_M)];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------- task building
// tpb[k] = ceil(cnt[k] / L); flags[0] = max cnt.  Buckets cut into several tasks are listed for the fold kernels
// (zkm_msm_curve.cuh) in three classes by their number t of partial sums, S = ZKM_FOLD_SEG:
//   A  2 <= t <= S      fold_list[i], i < flags[4]                     one quad sums them in place
//   M  S < t <= S^2     fold_list[K + 1 + i], i < flags[7]             segment sums, then one quad over the <= S segments
//   L  t > S^2          fold_list[K - 1 - i], i < flags[5]             segment sums, then one CTA per bucket
// Segments (S consecutive partial sums) of M and L buckets: seg_first[k] = first segment of bucket k, segtab[2 s] = k,
// segtab[2 s + 1] = segment number inside the bucket, flags[6] = segments in all.
__global__ void k_tasks_count(const uint32_t* __restrict__ cnt, uint32_t K, uint32_t L, uint32_t* __restrict__ tpb,
                              uint32_t* __restrict__ flags, uint32_t* __restrict__ fold_list, uint32_t* __restrict__ seg_first,
                              uint32_t* __restrict__ segtab) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = 0;
    if (k < K) {
        v = cnt[k];
        const uint32_t t = (v + L - 1) / L;
        tpb[k] = t;
        if (t > 1) {
            if (t <= ZKM_FOLD_SEG) {
                fold_list[atomicAdd(&flags[4], 1u)] = k;
            } else {
                const uint32_t nseg = (t + ZKM_FOLD_SEG - 1) / ZKM_FOLD_SEG;
                const uint32_t first = atomicAdd(&flags[6], nseg);
                if (nseg <= ZKM_FOLD_SEG) fold_list[K + 1 + atomicAdd(&flags[7], 1u)] = k;
                else fold_list[K - 1 - atomicAdd(&flags[5], 1u)] = k;
                seg_first[k] = first;
                for (uint32_t j = 0; j < nseg; j++) {
                    segtab[2 * (first + j)] = k;
                    segtab[2 * (first + j) + 1] = j;
                }
            }
        }
    } else if (k == K) {
        tpb[k] = 0;
    }
    // block max then one atomic
    __shared__ uint32_t smax;
    if (threadIdx.x == 0) smax = 0;
    __syncthreads();
    uint32_t wmax = __reduce_max_sync(0xffffffffu, v);
    if ((threadIdx.x & 31) == 0 && wmax) atomicMax(&smax, wmax);
    __syncthreads();
    if (threadIdx.x == 0 && smax) atomicMax(&flags[0], smax);
}

// per task: owner bucket by binary search in tbase, (start, len), histogram of lengths
__global__ void k_tasks_emit(const uint32_t* __restrict__ tbase, uint32_t K, const uint32_t* __restrict__ off,
                             const uint32_t* __restrict__ cnt, uint32_t L, uint32_t* __restrict__ tstart,
                             uint32_t* __restrict__ tlen, uint32_t* __restrict__ lenhist) {
    extern __shared__ uint32_t sh_hist[];  // L + 1
    for (uint32_t i = threadIdx.x; i <= L; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    const uint32_t T = tbase[K];
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
        // largest k with tbase[k] <= t
        uint32_t lo = 0, hi = K;  // tbase[lo] <= t < tbase[hi]
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (tbase[mid] <= t) lo = mid; else hi = mid;
        }
        uint32_t j = t - tbase[lo];
        uint32_t c = cnt[lo];
        uint32_t len = c - j * L;
        if (len > L) len = L;
        tstart[t] = off[lo] + j * L;
        tlen[t] = len;
        atomicAdd(&sh_hist[len], 1u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i <= L; i += blockDim.x)
        if (sh_hist[i]) atomicAdd(&lenhist[i], sh_hist[i]);
}

// cursor[len] = number of tasks strictly longer than len (descending-length order)
__global__ void k_len_offsets(const uint32_t* __restrict__ lenhist, uint32_t L, uint32_t* __restrict__ cursor) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t acc = 0;
    for (int len = (int)L; len >= 0; len--) {
        cursor[len] = acc;
        acc += lenhist[len];
    }
}

__global__ void k_tasks_order(const uint32_t* __restrict__ tlen, const uint32_t* __restrict__ tbase, uint32_t K,
                              uint32_t* __restrict__ cursor, uint32_t* __restrict__ order) {
    const uint32_t T = tbase[K];
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t t0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; t0 < T; t0 += gridDim.x * blockDim.x) {
        uint32_t t = t0 + lane;
        bool valid = t < T;
        uint32_t len = valid ? tlen[t] : 0xffffffffu;
        // warp-aggregated: lanes with equal len share one atomic
        uint32_t peers = __match_any_sync(0xffffffffu, len);
        uint32_t leader = __ffs(peers) - 1;
        uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (valid && lane == leader) base = atomicAdd(&cursor[len], (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (valid) order[base + rank] = t;
    }
}

// ---------------------------------------------------------------------------------- host side
static int windows_for(int scalar_bits, int c) { return (scalar_bits + 1 + c - 1) / c; }

// precomputed != 0: the windows share one bucket set (bases registered with window multiples)
//
// Cost model in units of one field product, fitted to sweeps on B200 (profiles/experiment_window_*.jsonl):
//   * per (point, window): 11 products while the XYZZ accumulation does the work (the 10 of a mixed addition plus the
//     share of sort and folds), 9.3 falling to 8.3 once the batched-affine levels take over (>= 24 M entries);
//   * per bucket: 40 (running-sum reduction, ~1.7 ns at BLS12-381 G1) plus a latency term that grows with log2 of the
//     bucket count -- a reduction over 0.5 M buckets takes 3.8 ms where the throughput term alone says 1 ms;
//   * SKEW of the top window: it only holds top = bits - c (W - 1) real scalar bits, so its n entries pile into 2^top
//     buckets: serialised atomics in the histogram / scatter kernels, long lists, extra fold levels.  Measured at
//     2^23: c = 17 (top 0), 18 (top 3), 19 (top 8) are 7-9 ms slower than the model without this term, c = 16 and 20
//     (top 15) are not -- which is why c = 16 wins from 2^19 to 2^23 although it has the most windows.
static int auto_window_bits(int curve, int group, size_t n, int precomputed) {
    const int bits = fr_bits(curve);
    if (n < 2) n = 2;
    const double limbs = coord_words(curve, 1) * 2.0;
    const double mads = (2.0 * limbs * limbs + limbs) * ((group == 2 && curve != ZKM_CURVE_BW6_761) ? 3.0 : 1.0);
    double best = 1e300;
    int best_c = 4;
    for (int c = 4; c <= (precomputed ? 24 : 21); c++) {
        double W = windows_for(bits, c);
        double B = (double)(1u << (c - 1));
        double cost;
        if (precomputed) {
            cost = W * (double)n * 10.0 + B * 50.0;
        } else {
            const double K = W * B;
            int top = bits - c * ((int)W - 1);
            if (top < 0) top = 0;
            if (top > c - 1) top = c - 1;
            const double E = W * (double)n;
            double per_entry = 11.0;
            if (E >= 24.0e6) {
                double lgE = log2(E / 24.0e6) * 0.5;
                per_entry = 9.3 - (lgE < 1.0 ? lgE : 1.0);
            }
            cost = E * per_entry + K * 40.0;
            if (K > 32768.0) {
                double lg = log2(K / 32768.0);
                cost += 9.0e6 * (lg < 6.0 ? lg : 6.0);
            }
            cost += 30.0 * (double)n * (1.0 - (double)top / (double)(c - 1));
        }
        if (cost < best) {
            best = cost;
            best_c = c;
        }
    }
    if (precomputed && n < ((size_t)1 << 21)) {
        // small registered MSMs are latency-bound: the serial chains of the single-bucket-set reduction grow
        // with log2(B) = c - 1, so a smaller window wins although it adds entries (measured on B200, witness
        // scalars: 2^16 G2 6.4 ms at c = 16 vs 4.0 ms at c = 9; 2^18 best at 12; 2^20 at 15-16)
        int lg = 0;
        while (((size_t)1 << (lg + 1)) <= n) lg++;
        int c_lat = (3 * lg) / 2 - 15;
        if (c_lat < 8) c_lat = 8;
        if (c_lat < best_c) best_c = c_lat;
    }
    return best_c;
}

int msm_auto_window_bits(int curve, int group, size_t n) {
    return auto_window_bits(curve, group, n, 0);
}

static const CurveOps* curve_ops(int curve, int group) {
    if (curve == ZKM_CURVE_BLS12_381 && group == 1) return ops_g1_bls();
    if (curve == ZKM_CURVE_BLS12_381 && group == 2) return ops_g2_bls();
    if (curve == ZKM_CURVE_BN254 && group == 1) return ops_g1_bn();
    if (curve == ZKM_CURVE_BN254 && group == 2) return ops_g2_bn();
    if (curve == ZKM_CURVE_BW6_761 && group == 1) return ops_g1_bw6();
    if (curve == ZKM_CURVE_BW6_761 && group == 2) return ops_g2_bw6();
    ZKM_FAIL(ZKM_ERR_ARG, "unknown curve %d / group %d", curve, group);
}

enum WsSlot {
    WS_COUNTS = 0, WS_OFF, WS_CURSOR, WS_IDX, WS_TPB_A, WS_TBASE_A, WS_FOLDLIST, WS_FOLDSEG, WS_TSTART, WS_TLEN,
    WS_ORDER, WS_LENHIST, WS_LENCUR, WS_PART_A, WS_FOLDSTAGE, WS_CONTRIB, WS_WSUM, WS_FLAGS, WS_CUBTMP,
    WS_AOFF_A, WS_AOFF_B, WS_ALEN_A, WS_ALEN_B, WS_PT_A, WS_PT_B, WS_PRE, WS_T, WS_PRE2, WS_XARR, WS_PAIRMAP, WS_DIGITS
};

static void exclusive_scan(Context* c, const uint32_t* in, uint32_t* out, size_t count, cudaStream_t s) {
    size_t tmp_bytes = 0;
    ZKM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, in, out, (int)count, s));
    void* tmp = c->ws[WS_CUBTMP].get(tmp_bytes ? tmp_bytes : 16);
    ZKM_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, (int)count, s));
}

// Window multiples for one part of a registration: picks the window size once (it is then fixed for every
// MSM over this part) and fills part->d_table.
void msm_precompute(Context* c, int curve, int group, BasesPart* part, cudaStream_t s) {
    const CurveOps* ops = curve_ops(curve, group);
    if (part->n == 0) return;
    int cb = c->opt.msm_window_bits > 0 ? c->opt.msm_window_bits : auto_window_bits(curve, group, part->n, 1);
    if (cb < 2) cb = 2;
    if (cb > 24) cb = 24;
    int W = windows_for(ops->scalar_bits, cb);
    if ((double)part->n * W >= 2.0e9) ZKM_FAIL(ZKM_ERR_ARG, "precomputed table of %zu x %d points exceeds 2^31 entries", part->n, W);
    const size_t rec = 2 * (size_t)coord_words(curve, group) * 8;
    ZKM_CUDA(malloc_retry((void**)&part->d_table, part->n * (size_t)W * rec));
    ops->precompute(s, part->d_xy, part->d_inf, (uint64_t)part->n, cb, W, part->d_table);
    part->pre_c = cb;
    part->pre_W = W;
}

void msm_run(Context* c, int curve, int group, const void* d_bases, const uint8_t* d_inf, const uint64_t* d_scalars,
             size_t n, uint64_t* d_out, cudaStream_t s, const BasesPart* pre, size_t pre_offset, const ScalarFeed* feed) {
    const CurveOps* ops = curve_ops(curve, group);
    if (n == 0) {
        ops->write_identity(s, d_out);
        return;
    }
    if (n >= (1ull << 31)) ZKM_FAIL(ZKM_ERR_ARG, "MSM of %zu points: at most 2^31 - 1 supported", n);
    MsmPlan pl;
    pl.scalar_bits = ops->scalar_bits;
    const bool use_pre = pre && pre->d_table;
    pl.c = use_pre ? pre->pre_c : (c->opt.msm_window_bits > 0 ? c->opt.msm_window_bits : msm_auto_window_bits(curve, group, n));
    if (pl.c < 2) pl.c = 2;
    if (pl.c > 24) pl.c = 24;
    pl.W = windows_for(pl.scalar_bits, pl.c);
    pl.B = 1u << (pl.c - 1);
    if (use_pre) {          // one bucket set, entries index the table of window multiples
        pl.K = pl.B;
        pl.key_stride = 0;
        pl.idx_stride = (uint32_t)pre->n;
        pl.idx_base = (uint32_t)pre_offset;
        pl.RW = 1;
        d_bases = pre->d_table;
    } else {
        pl.K = (uint32_t)pl.W * pl.B;
        pl.key_stride = pl.B;
        pl.idx_stride = 0;
        pl.idx_base = 0;
        pl.RW = pl.W;
    }
    if ((double)n * pl.W >= 4.0e9)
        ZKM_FAIL(ZKM_ERR_ARG, "MSM of %zu points x %d windows exceeds 2^32 bucket entries", n, pl.W);
    const uint32_t K = pl.K;
    const size_t entries = n * (size_t)pl.W;

    // chunk lengths: level 1 (affine gather) and the fold levels (partial sums)
    uint32_t L1;
    if (c->opt.msm_chunk > 0) {
        L1 = (uint32_t)c->opt.msm_chunk;
    } else {
        double per_thread = (double)entries / ((double)c->sm_count * 256.0 * 4.0);
        L1 = per_thread < 16 ? 16u : (per_thread > 256 ? 256u : (uint32_t)per_thread);
    }
    if (L1 > 1024) L1 = 1024;
    const size_t T1max = entries / L1 + K + 1;
    const size_t XB = ops->xyzz_bytes;

    uint32_t* counts = c->ws[WS_COUNTS].as<uint32_t>(K + 1);
    uint32_t* off = c->ws[WS_OFF].as<uint32_t>(K + 1);
    uint32_t* cursor = c->ws[WS_CURSOR].as<uint32_t>(K + 1);
    uint32_t* idx = c->ws[WS_IDX].as<uint32_t>(entries);
    uint32_t* tpb = c->ws[WS_TPB_A].as<uint32_t>(K + 1);
    uint32_t* tbase = c->ws[WS_TBASE_A].as<uint32_t>(K + 1);
    uint32_t* fold_list = c->ws[WS_FOLDLIST].as<uint32_t>(2 * ((size_t)K + 1));
    // segments of the long buckets: a bucket is long with > ZKM_FOLD_SEG tasks and its last segment may be short
    const size_t max_segs = T1max / ZKM_FOLD_SEG + T1max / (ZKM_FOLD_SEG + 1) + 2;
    uint32_t* seg_first = c->ws[WS_FOLDSEG].as<uint32_t>(K + 1 + 2 * max_segs);
    uint32_t* segtab = seg_first + K + 1;
    uint32_t* tstart = c->ws[WS_TSTART].as<uint32_t>(T1max);
    uint32_t* tlen = c->ws[WS_TLEN].as<uint32_t>(T1max);
    uint32_t* order = c->ws[WS_ORDER].as<uint32_t>(T1max);
    uint32_t* lenhist = c->ws[WS_LENHIST].as<uint32_t>(1024 + 2);
    uint32_t* lencur = c->ws[WS_LENCUR].as<uint32_t>(1024 + 2);
    char* part = (char*)c->ws[WS_PART_A].get(T1max * XB);
    char* fold_stage = (char*)c->ws[WS_FOLDSTAGE].get(max_segs * XB);
    // device words: [0] largest list, [1] a scalar has bits above the modulus width, [2] XYZZ tasks (profile),
    // [4] / [7] / [5] fold classes A / M / L, [6] segments.  Nothing here is read by the host during the run.
    uint32_t* flags = c->ws[WS_FLAGS].as<uint32_t>(8);

    const bool prof = c->opt.profile != 0;
    // work counters of a profiled run (zkm_profile_last_msm_counts): device words read back after the run
    const uint32_t* cnt_words[16] = {nullptr};
    int n_cnt_levels = 0;
    auto mark = [&](int i) {
        if (!prof) return;
        if (!c->pev[i]) ZKM_CUDA(cudaEventCreate(&c->pev[i]));
        ZKM_CUDA(cudaEventRecord(c->pev[i], s));
    };
    c->pev_valid = false;
    mark(0);
    const unsigned grid_stream = (unsigned)c->sm_count * 8;
    ZKM_CUDA(cudaMemsetAsync(counts, 0, (K + 1) * sizeof(uint32_t), s));
    ZKM_CUDA(cudaMemsetAsync(flags, 0, 8 * sizeof(uint32_t), s));
    // ---- how many batched-affine levels this run takes (needed first: the x-coordinate array of level 0 depends only on
    // the bases, so it is built BEFORE the histogram pass waits for the scalars -- behind a host upload it costs nothing)
    int n_aff = c->opt.msm_affine_levels;
    if (n_aff < 0) {
        // Measured on B200 (profiles/experiment_affine_threshold_r1p.jsonl, BLS12-381 G1): a pairwise level costs
        // ~0.33 ns per pair plus ~1 ms of fixed latency (inversion kernel, scans), the XYZZ accumulation it saves
        // ~0.42 ns per entry -- so a level pays while it still has more than ~6 M pairs, and the whole scheme from
        // ~24 M entries on (2^21 points: 18.9 -> 16.6 ms with two levels; 2^20: no gain).  Cheaper field products
        // (BN254: 136 wide MADs instead of 300) raise both thresholds in proportion; they are NOT lowered for the
        // heavier fields (G2, BW6-761), whose pair kernels run at 2 CTAs/SM and showed no gain below ~24 M entries
        // (profiles/experiment_affine_rule_r1r.jsonl).
        const double limbs = (double)ops->coord_bytes / 4.0 / (group == 2 && curve != ZKM_CURVE_BW6_761 ? 2.0 : 1.0);
        const double mads = (2.0 * limbs * limbs + limbs) * (group == 2 && curve != ZKM_CURVE_BW6_761 ? 3.0 : 1.0);
        const double scale = mads < 300.0 ? 300.0 / mads : 1.0;   // 1 for BLS12-381 G1
        n_aff = 0;
        if ((double)entries >= 24.0e6 * scale && entries < ((size_t)1 << 31)) {   // map words keep bit 31 for the pair flag
            double avg = (double)entries / (double)K;   // a level needs lists of >= 4 entries on average
            double pairs = (double)entries * 0.5;
            while (n_aff < 8 && avg >= 4.0 && pairs > 6.0e6 * scale) {
                avg *= 0.5;
                pairs *= 0.5;
                n_aff++;
            }
        }
    }
    if (entries >= ((size_t)1 << 31)) n_aff = 0;   // (also when the level count was forced by option)
    // x coordinates of the bases in 64-byte slots for the level-0 forward gathers (one DRAM burst per x);
    // rebuilt per call (2.5 GB of streaming traffic at 2^24, ~0.4 ms) -- not for tables of window multiples
    const void* xarr = nullptr;
    if (n_aff > 0 && ops->xarr_slot > 0 && !use_pre && c->opt.msm_xarr) {
        void* xa = c->ws[WS_XARR].get(n * (size_t)ops->xarr_slot);
        ops->build_xarr((unsigned)c->sm_count, s, d_bases, (uint64_t)n, xa);
        xarr = xa;
    }
    const bool wide = fr_words(curve) == 6;   // 377-bit scalars (BW6-761)
    // all windows scattered by one launch when the whole list array fits in L2; otherwise one launch per window
    // (write set n x 4 B inside L2), fed from the digit words the histogram pass leaves behind
    const bool one_scatter = entries * sizeof(uint32_t) <= (96u << 20);
    uint32_t* dig = one_scatter ? nullptr : c->ws[WS_DIGITS].as<uint32_t>(entries);
    {
        // the histogram pass: all scalars at once, or chunk by chunk behind the caller's copies (ScalarFeed)
        ScalarFeed whole;
        whole.n = 1;
        whole.begin[0] = 0;
        whole.begin[1] = n;
        const ScalarFeed& f = (feed && feed->n > 0) ? *feed : whole;
        for (int ch = 0; ch < f.n; ch++) {
            if (f.ev[ch]) ZKM_CUDA(cudaStreamWaitEvent(s, f.ev[ch], 0));
            const uint64_t b = f.begin[ch], e = f.begin[ch + 1];
            if (e <= b) continue;
            const uint64_t need = (e - b + 255) / 256;
            const unsigned grid = (unsigned)(need < grid_stream ? need : grid_stream);
            if (wide)
                ZKM_LAUNCH((k_msm_digits<0, 12>), grid, 256, 0, s, (const uint32_t*)d_scalars, d_inf, (uint64_t)n, pl, -1, counts,
                           (uint32_t*)nullptr, flags, dig, b, e);
            else
                ZKM_LAUNCH((k_msm_digits<0, 8>), grid, 256, 0, s, (const uint32_t*)d_scalars, d_inf, (uint64_t)n, pl, -1, counts,
                           (uint32_t*)nullptr, flags, dig, b, e);
        }
    }
    exclusive_scan(c, counts, off, K + 1, s);
    ZKM_CUDA(cudaMemcpyAsync(cursor, off, K * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    if (one_scatter) {
        if (wide)
            ZKM_LAUNCH((k_msm_digits<1, 12>), grid_stream, 256, 0, s, (const uint32_t*)d_scalars, d_inf, (uint64_t)n, pl, -1,
                       cursor, idx, flags, (uint32_t*)nullptr, (uint64_t)0, (uint64_t)n);
        else
            ZKM_LAUNCH((k_msm_digits<1, 8>), grid_stream, 256, 0, s, (const uint32_t*)d_scalars, d_inf, (uint64_t)n, pl, -1,
                       cursor, idx, flags, (uint32_t*)nullptr, (uint64_t)0, (uint64_t)n);
    } else {
        for (int w = 0; w < pl.W; w++)
            ZKM_LAUNCH(k_msm_scatter_row, grid_stream, 256, 0, s, (const uint32_t*)(dig + (size_t)w * n), (uint64_t)n, pl, w, cursor, idx);
    }

    mark(1);
    const unsigned kblocks = (K + 1 + 255) / 256;
    // ---- batched-affine pairwise levels (large MSMs): halve every bucket list n_aff times
    const uint32_t* cur_off = off;       // offsets / lengths / source of the current bucket lists
    const uint32_t* cur_cnt = counts;
    const void* cur_src = d_bases;
    const uint32_t* cur_idx = idx;
    {
        const size_t CBy = ops->coord_bytes;
        const uint32_t m = (uint32_t)c->opt.msm_pair_m, m2 = (uint32_t)c->opt.msm_pair_m2;
        size_t Eb = entries;
        int p = 0;
        for (int lvl = 0; lvl < n_aff; lvl++) {
            const size_t Eout = (Eb + (K < Eb ? K : Eb)) / 2 + 1;      // bound on the outputs of this level
            const size_t nT = 32 * ((Eout + 32 * (size_t)m - 1) / (32 * (size_t)m)), nU = (nT + m2 - 1) / m2;   // chains (pair_chains)
            uint32_t* aoff = c->ws[p ? WS_AOFF_B : WS_AOFF_A].as<uint32_t>(K + 1);
            uint32_t* alen = c->ws[p ? WS_ALEN_B : WS_ALEN_A].as<uint32_t>(K + 1);
            char* pt = (char*)c->ws[p ? WS_PT_B : WS_PT_A].get(Eout * 2 * CBy);
            char* pre = (char*)c->ws[WS_PRE].get(Eout * CBy);
            char* Tt = (char*)c->ws[WS_T].get((nT + 1) * CBy);
            char* pre2 = (char*)c->ws[WS_PRE2].get((nT + 1) * CBy);
            uint32_t* pmap = c->ws[WS_PAIRMAP].as<uint32_t>(Eout);
            ZKM_LAUNCH(k_pair_lens, kblocks, 256, 0, s, cur_off, K, alen);
            exclusive_scan(c, alen, aoff, K + 1, s);
            {
                const uint64_t need = Eout / (256 * ZKM_MAP_M) + 1;
                ZKM_LAUNCH(k_pair_map, (unsigned)(need < grid_stream ? need : grid_stream), 256, 0, s, cur_off, aoff, K, pmap);
            }
            ops->pair_fwd((unsigned)c->sm_count, nT, s, lvl == 0, cur_src, cur_idx, pmap, aoff, K, m, pre, Tt, lvl == 0 ? xarr : nullptr);
            ops->pair_inv((unsigned)c->sm_count, nU, s, aoff, K, m, m2, Tt, pre2);
            ops->pair_bwd((unsigned)c->sm_count, nT, s, lvl == 0, cur_src, cur_idx, pmap, aoff, K, m, pre, Tt, pt);
            if (n_cnt_levels < 8) cnt_words[4 + n_cnt_levels++] = aoff + K;   // outputs of this level
            cur_off = aoff;
            cur_cnt = alen;
            cur_src = pt;
            cur_idx = nullptr;
            Eb = Eout;
            p ^= 1;
        }
    }
    mark(2);
    // Task list: every bucket list is cut into tasks of <= L1 entries, tasks sorted by length.  Buckets that were cut
    // (a long list: the skewed top window, the "ones" bucket of a Groth16 witness) get their partial sums folded by the
    // two fold kernels below -- a fixed launch sequence: the run needs NO device read-back, the host thread never waits
    // (round 1/2a: log_f levels of {count, scan, emit, order, accumulate} sized from a read-back of the largest list;
    // 6 levels = 48 launches and ~0.5 of the 3.4 ms of a 2^16-point MSM).
    ZKM_LAUNCH(k_tasks_count, kblocks, 256, 0, s, cur_cnt, K, L1, tpb, flags, fold_list, seg_first, segtab);
    exclusive_scan(c, tpb, tbase, K + 1, s);
    if (prof) ZKM_CUDA(cudaMemcpyAsync(flags + 2, tbase + K, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));  // task count
    ZKM_CUDA(cudaMemsetAsync(lenhist, 0, (L1 + 1) * sizeof(uint32_t), s));
    ZKM_LAUNCH(k_tasks_emit, grid_stream, 256, (L1 + 1) * sizeof(uint32_t), s, tbase, K, cur_off, cur_cnt, L1, tstart, tlen,
               lenhist);
    ZKM_LAUNCH(k_len_offsets, 1, 32, 0, s, lenhist, L1, lencur);
    ZKM_LAUNCH(k_tasks_order, grid_stream, 256, 0, s, tlen, tbase, K, lencur, order);
    mark(3);
    // One task per thread, persistent over the task list; the accumulators are register-bound (two 128-thread CTAs per
    // SM).  A small MSM gets only the CTAs its task bound needs: one warp per scheduler (a dependent product is 1.85 k
    // cycles alone, 2.65 k with two warps per scheduler: tools/microbench/tail_latency.cu) and room on the SMs for the
    // kernels of concurrent MSMs (CTAs beyond the device-side task count exit at once).
    {
        const size_t need = (T1max + 127) / 128, cap = (size_t)c->sm_count * (size_t)ops->accum_ctas_per_sm;
        ops->accum_affine((unsigned)(need < cap ? need : cap), s, cur_src, cur_idx, TaskList{tstart, tlen, order, tbase, K}, part);
    }
    mark(4);
    ops->fold((unsigned)c->sm_count, s, part, tbase, tpb, fold_list, seg_first, segtab, fold_stage, K, (uint32_t)max_segs, flags + 4);
    mark(5);
    void* contrib = c->ws[WS_CONTRIB].get(msm_contrib_records(pl.RW, pl.B) * XB);
    void* wsum = c->ws[WS_WSUM].get((size_t)pl.RW * XB);
    ops->reduce(s, part, tbase, tpb, pl, contrib, wsum, flags, d_out);
    mark(6);
    c->pev_valid = prof;
    if (prof) {
        // [0] points, [1] windows, [2] window bits, [3] list entries (non-zero digits), [4..4+L) outputs of each
        // batched-affine level, [12] affine levels L, [13] XYZZ tasks, [14] buckets K, [15] buckets folded
        for (auto& v : c->pcount) v = 0;
        cnt_words[3] = off + K;
        cnt_words[13] = flags + 2;
        ZKM_CUDA(cudaStreamSynchronize(s));
        for (int i = 0; i < 16; i++) {
            if (!cnt_words[i]) continue;
            uint32_t w = 0;
            ZKM_CUDA(cudaMemcpy(&w, cnt_words[i], sizeof(uint32_t), cudaMemcpyDeviceToHost));
            c->pcount[i] = w;
        }
        c->pcount[0] = n;
        c->pcount[1] = (uint64_t)pl.W;
        c->pcount[2] = (uint64_t)pl.c;
        c->pcount[12] = (uint64_t)n_cnt_levels;
        c->pcount[14] = K;
        {   // buckets whose partial sums were folded (quad kernel + CTA kernel)
            uint32_t nf[4] = {0, 0, 0, 0};
            ZKM_CUDA(cudaMemcpy(nf, flags + 4, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
            c->pcount[15] = (uint64_t)nf[0] + nf[1] + nf[3];
        }
        note_profiled_lane(c);
    }
}

void points_sum_run(Context* c, int curve, int group, const uint64_t* d_points, size_t m, uint64_t* d_out,
                    cudaStream_t s) {
    (void)c;
    curve_ops(curve, group)->points_sum(s, d_points, (uint64_t)m, d_out);
}

void testgen_progression(Context* c, int curve, int group, uint64_t a0, uint64_t d, size_t n, uint64_t* d_out,
                         cudaStream_t s) {
    (void)c;
    curve_ops(curve, group)->gen_progression(s, a0, d, (uint64_t)n, d_out);
}

}  // namespace zkm

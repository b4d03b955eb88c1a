// zkm_msm.cuh -- interface between the curve-independent MSM driver (zkm_msm.cu: digit extraction,
// bucket sort, task lists) and the per-curve translation units (zkm_msm_g{1,2}_{bls,bn}.cu: bucket
// accumulation and window reduction, instantiated from zkm_msm_curve.cuh).  Splitting them keeps
// every ptxas job small enough to build in parallel.
#pragma once
#include "zkm_common.cuh"

namespace zkm {

struct MsmPlan {
    int c;         // window bits
    int W;         // windows = ceil((scalar_bits + 1) / c)
    uint32_t B;    // buckets per window = 2^(c-1)  (signed digits: |d| in 1..B)
    uint32_t K;    // number of bucket lists: W * B, or B when the windows share one bucket set
    int scalar_bits;
    // bucket key = w * key_stride + |d| - 1 ; list entry = w * idx_stride + idx_base + i  (| sign << 31)
    // plain bases: key_stride = B, idx_stride = 0.  Bases registered with precomputed window multiples
    // 2^(c w) P_i (table row w): key_stride = 0 (one bucket set for all windows), idx_stride = table row length.
    uint32_t key_stride;
    uint32_t idx_stride;
    uint32_t idx_base;
    int RW;        // windows seen by the reduction: W, or 1 with precomputed multiples (no Horner doublings)
};

// Task list of one fold level: task t sums entries [tstart[t], tstart[t] + tlen[t]); threads walk
// `order` (tasks sorted by decreasing length) so that the lanes of a warp run equally long loops.
struct TaskList {
    const uint32_t* tstart;
    const uint32_t* tlen;
    const uint32_t* order;
    const uint32_t* tbase;  // tbase[K] = number of tasks (device-resident, never read by the host)
    uint32_t K;
};

// partial sums per fold segment: a bucket with up to this many is summed by one quad, longer ones segment by segment
constexpr uint32_t ZKM_FOLD_SEG = 8;

struct CurveOps {
    int curve, group, scalar_bits;
    size_t xyzz_bytes;
    // K4: out[task] = sum of the (sign-adjusted) affine bases named by idx[...]
    void (*accum_affine)(unsigned grid, cudaStream_t s, const void* bases, const uint32_t* idx, TaskList tl, void* out);
    int accum_ctas_per_sm;   // resident 128-thread CTAs of the accumulation kernel (register-bound)
    // fold: bucket k holds tpb[k] partial sums items[tbase[k] ..); afterwards its sum is items[tbase[k]].  fold_list,
    // seg_first, segtab, n_lists as written by k_tasks_count; `stage` holds one record per segment.
    void (*fold)(unsigned sm_count, cudaStream_t s, void* items, const uint32_t* tbase, const uint32_t* tpb,
                 const uint32_t* fold_list, const uint32_t* seg_first, const uint32_t* segtab, void* stage, uint32_t K,
                 uint32_t max_segs, const uint32_t* n_lists);
    // K5: bucket sums (cnt[k] != 0: item at off[k]) -> affine result record at d_out; flags[1] != 0 (a scalar was not
    // canonical) turns the record's flag word into 2
    void (*reduce)(cudaStream_t s, const void* items, const uint32_t* off, const uint32_t* cnt, MsmPlan pl, void* contrib,
                   void* wsum, const uint32_t* flags, uint64_t* d_out);
    void (*write_identity)(cudaStream_t s, uint64_t* d_out);
    void (*points_sum)(cudaStream_t s, const uint64_t* d_points, uint64_t m, uint64_t* d_out);
    void (*gen_progression)(cudaStream_t s, uint64_t a0, uint64_t d, uint64_t n, void* d_out);
    // table[w * n + i] = affine(2^(c w) * bases[i]), w < W  (window multiples of registered bases)
    void (*precompute)(cudaStream_t s, const void* bases, const uint8_t* inf, uint64_t n, int c, int W, void* table);
    // batched-affine pairwise level (zkm_msm_affine.cuh): denominators + prefix products, inversion of the
    // per-thread totals, unwind + affine additions into the next level's array
    void (*pair_fwd)(unsigned sm_count, uint64_t nT_bound, cudaStream_t s, int level0, const void* src, const uint32_t* idx,
                     const uint32_t* map, const uint32_t* off_out, uint32_t K, uint32_t m, void* pre, void* T, const void* xarr);
    // level-0 x-coordinate array (xarr_slot bytes per base; 0 = this group does not use one)
    void (*build_xarr)(unsigned sm_count, cudaStream_t s, const void* bases, uint64_t n, void* xarr);
    int xarr_slot;
    void (*pair_inv)(unsigned sm_count, uint64_t nU_bound, cudaStream_t s, const uint32_t* off_out, uint32_t K, uint32_t m,
                     uint32_t m2, void* T, void* pre2);
    void (*pair_bwd)(unsigned sm_count, uint64_t nT_bound, cudaStream_t s, int level0, const void* src, const uint32_t* idx,
                     const uint32_t* map, const uint32_t* off_out, uint32_t K, uint32_t m, const void* pre, const void* Tinv,
                     void* dst);
    size_t coord_bytes;
};

// buckets per reduction thread: large enough that the lo * run scalar multiple is a small overhead
// (and small enough that a small MSM still spreads over the SMs: the chain of 2 g additions is pure latency)
inline uint32_t msm_reduce_group(uint32_t B, uint32_t W) {
    uint32_t g = 64;
    while (g > 4 && (uint64_t)W * (B / g) < 32768) g >>= 1;
    return g < B ? g : B;
}
// XYZZ records needed by CurveOps::reduce for `contrib`
inline size_t msm_contrib_records(int W, uint32_t B) {
    uint32_t per_w = B / msm_reduce_group(B, (uint32_t)W);
    // quad path: acc + slice sums (256 per slice).  Hierarchical path: acc + run of level 0, the upper levels (each at
    // most 1/4 of the level below + 1 per window, acc and run: < per_w + 64), 17 per-level window sums, slice sums of the
    // plain sums (one per 256 records of every level, rounded up per level and window: < per_w / 128 + 34)
    return 2 * (size_t)W * per_w + (size_t)W * ((per_w + 255) / 256) + (size_t)W * (per_w + 64) + 17 * (size_t)W +
           (size_t)W * (per_w / 128 + 34) + 16;
}

const CurveOps* ops_g1_bls();
const CurveOps* ops_g2_bls();
const CurveOps* ops_g1_bn();
const CurveOps* ops_g2_bn();
const CurveOps* ops_g1_bw6();
const CurveOps* ops_g2_bw6();

}  // namespace zkm

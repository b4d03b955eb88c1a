// zkm_msm_bw6_pair.cu -- the batched-affine pair kernels (zkm_msm_affine.cuh) instantiated for the 761-bit Fq of
// BW6-761, in a translation unit of their own: ptxas then works on this half and on zkm_msm_bw6.cu (accumulation,
// reduction, tails) in parallel, which halves the longest leg of a from-scratch build.
#include "zkm_msm_curve.cuh"

namespace zkm {

void bw6_pair_fwd(unsigned sm_count, uint64_t nT_bound, cudaStream_t s, int level0, const void* src, const uint32_t* idx,
                  const uint32_t* map, const uint32_t* off_out, uint32_t K, uint32_t m, void* pre, void* T, const void* xarr) {
    OpsImpl<G1Bw6>::pair_fwd(sm_count, nT_bound, s, level0, src, idx, map, off_out, K, m, pre, T, xarr);
}
void bw6_pair_inv(unsigned sm_count, uint64_t nU_bound, cudaStream_t s, const uint32_t* off_out, uint32_t K, uint32_t m,
                  uint32_t m2, void* T, void* pre2) {
    OpsImpl<G1Bw6>::pair_inv(sm_count, nU_bound, s, off_out, K, m, m2, T, pre2);
}
void bw6_pair_bwd(unsigned sm_count, uint64_t nT_bound, cudaStream_t s, int level0, const void* src, const uint32_t* idx,
                  const uint32_t* map, const uint32_t* off_out, uint32_t K, uint32_t m, const void* pre, const void* Tinv,
                  void* dst) {
    OpsImpl<G1Bw6>::pair_bwd(sm_count, nT_bound, s, level0, src, idx, map, off_out, K, m, pre, Tinv, dst);
}
void bw6_build_xarr(unsigned sm_count, cudaStream_t s, const void* bases, uint64_t n, void* xarr) {
    OpsImpl<G1Bw6>::build_xarr(sm_count, s, bases, n, xarr);
}

}  // namespace zkm

// zkm_msm_bw6.cu -- MSM bucket kernels instantiated for BW6-761 (see zkm_msm_curve.cuh).  G1 and G2 are both
// curves over the 761-bit Fq (24 x 32-bit limbs), so one translation unit serves both groups.
#include "zkm_msm_curve.cuh"

namespace zkm {
const CurveOps* ops_g1_bw6() {
    static const CurveOps o = OpsImpl<G1Bw6>::make(ZKM_CURVE_BW6_761, 1);
    return &o;
}
const CurveOps* ops_g2_bw6() {
    static const CurveOps o = OpsImpl<G2Bw6>::make(ZKM_CURVE_BW6_761, 2);
    return &o;
}
}  // namespace zkm

// zkm_msm_g2_bls.cu -- MSM bucket kernels instantiated for one group (see zkm_msm_curve.cuh).
#include "zkm_msm_curve.cuh"

namespace zkm {
const CurveOps* ops_g2_bls() {
    static const CurveOps o = OpsImpl<G2Bls>::make(ZKM_CURVE_BLS12_381, 2);
    return &o;
}
}  // namespace zkm

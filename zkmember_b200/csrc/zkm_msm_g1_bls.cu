// zkm_msm_g1_bls.cu -- MSM bucket kernels instantiated for one group (see zkm_msm_curve.cuh).
#include "zkm_msm_curve.cuh"

namespace zkm {
const CurveOps* ops_g1_bls() {
    static const CurveOps o = OpsImpl<G1Bls>::make(ZKM_CURVE_BLS12_381, 1);
    return &o;
}
}  // namespace zkm

// zkm_msm_g1_bn.cu -- MSM bucket kernels instantiated for one group (see zkm_msm_curve.cuh).
#include "zkm_msm_curve.cuh"

namespace zkm {
const CurveOps* ops_g1_bn() {
    static const CurveOps o = OpsImpl<G1Bn>::make(ZKM_CURVE_BN254, 1);
    return &o;
}
}  // namespace zkm

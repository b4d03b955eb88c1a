// zkm_msm_g2_bn.cu -- MSM bucket kernels instantiated for one group (see zkm_msm_curve.cuh).
#include "zkm_msm_curve.cuh"

namespace zkm {
const CurveOps* ops_g2_bn() {
    static const CurveOps o = OpsImpl<G2Bn>::make(ZKM_CURVE_BN254, 2);
    return &o;
}
}  // namespace zkm

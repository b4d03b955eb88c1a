"""zkmember_b200 -- B200-native MSM / NTT backend for zkMember's Groth16 and Marlin provers.

Only the hot path lives here: `csrc/` (hand-written CUDA for sm_100a + the C ABI of
include/zkm_b200.h) and the host-side mirrors of the two arkworks interfaces it replaces:

    from zkmember_b200 import VariableBaseMSM, Radix2EvaluationDomain
"""
from ._lib import ZkmError, DomainError, init, shutdown, set_option, launch_count, load  # noqa: F401
from .msm import VariableBaseMSM, RegisteredBases, AffinePoint, msm_window_bits  # noqa: F401
from .domain import Radix2EvaluationDomain, GeneralEvaluationDomain  # noqa: F401

__all__ = [
    "VariableBaseMSM", "RegisteredBases", "AffinePoint", "Radix2EvaluationDomain", "GeneralEvaluationDomain",
    "ZkmError", "DomainError", "init", "shutdown", "set_option", "launch_count", "load", "msm_window_bits",
]

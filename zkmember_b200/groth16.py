"""Host-side mirror of the hot part of ark-groth16 0.3.0's prover.

`witness_map` replaces the FFT section of `R1CStoQAP::witness_map` (src/r1cs_to_qap.rs; reached from
/root/reference/benches/groth16.rs:115): the caller (unchanged Rust/host code) evaluates the R1CS
matrices on the assignment -- a[i] = <A_i, z>, b[i] = <B_i, z>, c[i] = <C_i, z>, a[nc + j] = z_j --
and this runs the seven NTTs and the pointwise step on the GPU without leaving HBM.

`ProvingKeyMSMs` holds the five query vectors of a Groth16 proving key registered once on the device
(`pk` is reused for every proof, benches/groth16.rs:107-115) and runs the five MSMs of `create_proof`
(src/prover.rs: h_query, l_query, a_query, b_g1_query on G1, b_g2_query on G2).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .msm import RegisteredBases, AffinePoint, _curve_id


def witness_map(a, b, c, curve="bls12_381") -> np.ndarray:
    """h = coset_ifft((coset_fft(ifft a) * coset_fft(ifft b) - coset_fft(ifft c)) / (g^n - 1)).
    a, b, c: (n, 4) uint64 Montgomery evaluations over the domain, n a power of two."""
    cid = _curve_id(curve)
    S = _lib.FR_WORDS[cid]
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, S)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, S)
    c = np.ascontiguousarray(c, dtype=np.uint64).reshape(-1, S)
    n = len(a)
    log_n = n.bit_length() - 1
    if n == 0 or (1 << log_n) != n or len(b) != n or len(c) != n:
        raise ValueError("a, b, c must have the same power-of-two length")
    h = np.zeros_like(a)
    p = lambda x: ctypes.c_void_p(x.ctypes.data)
    _lib.check(_lib.lib().zkm_witness_map(cid, p(a), p(b), p(c), log_n, p(h)))
    return h


class ProvingKeyMSMs:
    """The five MSM bases of a Groth16 proving key, resident in HBM."""

    def __init__(self, curve, h_query, l_query, a_query, b_g1_query, b_g2_query, infinity=None, precompute=True):
        cid = _curve_id(curve)
        infinity = infinity or {}
        if precompute:
            _lib.set_option("msm_precompute", 1)
        try:
            self.h = RegisteredBases(cid, 1, h_query, infinity.get("h"))
            self.l = RegisteredBases(cid, 1, l_query, infinity.get("l"))
            self.a = RegisteredBases(cid, 1, a_query, infinity.get("a"))
            self.b_g1 = RegisteredBases(cid, 1, b_g1_query, infinity.get("b_g1"))
            self.b_g2 = RegisteredBases(cid, 2, b_g2_query, infinity.get("b_g2"))
        finally:
            if precompute:
                _lib.set_option("msm_precompute", 0)

    def prove_msms(self, h_scalars, aux_scalars, full_scalars) -> dict:
        """h_acc, l_acc and the MSM parts of g_a, g1_b, g2_b (calculate_coeff's `acc`).  Upstream runs the
        five MSMs one after another; here they are issued from five host threads, each borrowing its own
        library lane, so the G2 MSM and the serial tails of the G1 ones overlap on the GPU."""
        from concurrent.futures import ThreadPoolExecutor
        jobs = {
            "h_acc": (self.h, h_scalars), "l_acc": (self.l, aux_scalars), "a_acc": (self.a, full_scalars),
            "b_g1_acc": (self.b_g1, full_scalars), "b_g2_acc": (self.b_g2, full_scalars),
        }
        with ThreadPoolExecutor(max_workers=5) as ex:
            futs = {k: ex.submit(reg.msm, sc) for k, (reg, sc) in jobs.items()}
            return {k: f.result() for k, f in futs.items()}

    def release(self):
        for r in (self.h, self.l, self.a, self.b_g1, self.b_g2):
            r.release()

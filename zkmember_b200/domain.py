"""Host-side mirror of ark-poly 0.3.0's Radix2EvaluationDomain.

Replaces `Radix2EvaluationDomain::{new, fft, ifft, coset_fft, coset_ifft}` (+ the `_in_place`
forms; src/domain/radix2/{mod,fft}.rs and the EvaluationDomain trait defaults in
src/domain/mod.rs; pin /root/reference/Cargo.lock:338-339), reached from ark-groth16's
witness_map (/root/reference/benches/groth16.rs:115) and Marlin's AHP prover
(benches/marlin.rs:202,311).  Same names, argument meaning and error behaviour as upstream:
`new(num_coeffs)` rounds up to a power of two and returns None when it exceeds the field's
two-adicity; `fft` takes coefficients (zero-padded to the domain size) and returns evaluations
in natural order; `ifft` the inverse, etc.  Field elements are numpy uint64 (n, 4) in Montgomery
form, exactly ark-ff's Fp256 limbs.  All arithmetic runs on the GPU through the C ABI.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib


def _curve_id(curve) -> int:
    return _lib.CURVE_IDS[curve] if isinstance(curve, str) else int(curve)


TWO_ADICITY = {_lib.CURVE_BLS12_381: 32, _lib.CURVE_BN254: 28, _lib.CURVE_BW6_761: 46}


class Radix2EvaluationDomain:
    def __init__(self, curve, log_size_of_group: int):
        self.curve = _curve_id(curve)
        self.log_size_of_group = int(log_size_of_group)
        self.size = 1 << self.log_size_of_group
        self.words = _lib.FR_WORDS[self.curve]
        consts = np.zeros((5, self.words), dtype=np.uint64)
        _lib.check(_lib.lib().zkm_domain_constants(self.curve, self.log_size_of_group,
                                                   ctypes.c_void_p(consts.ctypes.data)))
        self.group_gen, self.group_gen_inv, self.size_inv, self.generator, self.generator_inv = (
            consts[i].copy() for i in range(5))

    @classmethod
    def new(cls, num_coeffs: int, curve="bls12_381"):
        """Radix2EvaluationDomain::new: size = num_coeffs.next_power_of_two(); None if the field has
        no subgroup of that order."""
        n = max(int(num_coeffs), 1)
        log_n = (n - 1).bit_length()
        if log_n > TWO_ADICITY[_curve_id(curve)]:
            return None
        return cls(curve, log_n)

    # -- helpers
    def _prepare(self, x) -> np.ndarray:
        a = np.array(x, dtype=np.uint64, order="C").reshape(-1, self.words)
        if len(a) > self.size:
            raise ValueError("input of %d elements exceeds the domain size %d" % (len(a), self.size))
        if len(a) < self.size:  # upstream: coeffs.resize(self.size(), zero)
            a = np.concatenate([a, np.zeros((self.size - len(a), self.words), dtype=np.uint64)])
        return a

    def _run_in_place(self, a: np.ndarray, inverse: bool, coset: bool):
        if a.dtype != np.uint64 or not a.flags["C_CONTIGUOUS"] or a.size != self.words * self.size:
            raise ValueError("in-place transforms need a C-contiguous uint64 array of exactly size x %d words" % self.words)
        _lib.check(_lib.lib().zkm_ntt(self.curve, ctypes.c_void_p(a.ctypes.data), self.log_size_of_group,
                                      int(inverse), int(coset)))

    # -- the EvaluationDomain surface
    def fft_in_place(self, coeffs: np.ndarray):
        self._run_in_place(coeffs, False, False)

    def ifft_in_place(self, evals: np.ndarray):
        self._run_in_place(evals, True, False)

    def coset_fft_in_place(self, coeffs: np.ndarray):
        self._run_in_place(coeffs, False, True)

    def coset_ifft_in_place(self, evals: np.ndarray):
        self._run_in_place(evals, True, True)

    def fft(self, coeffs) -> np.ndarray:
        a = self._prepare(coeffs)
        self.fft_in_place(a)
        return a

    def ifft(self, evals) -> np.ndarray:
        a = self._prepare(evals)
        self.ifft_in_place(a)
        return a

    def coset_fft(self, coeffs) -> np.ndarray:
        a = self._prepare(coeffs)
        self.coset_fft_in_place(a)
        return a

    def coset_ifft(self, evals) -> np.ndarray:
        a = self._prepare(evals)
        self.coset_ifft_in_place(a)
        return a

    # -- device-resident variant (pointers into HBM, e.g. torch tensors' data_ptr())
    def transform_device(self, d_in: int, d_out: int, inverse: bool = False, coset: bool = False, stream: int = 0):
        _lib.check(_lib.lib().zkm_ntt_device(self.curve, ctypes.c_void_p(d_in), ctypes.c_void_p(d_out),
                                             self.log_size_of_group, int(inverse), int(coset),
                                             ctypes.c_void_p(stream)))


GeneralEvaluationDomain = Radix2EvaluationDomain  # BLS12-381 / BN254 Fr have no mixed-radix parameters

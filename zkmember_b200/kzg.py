"""Host-side mirror of the hot part of ark-poly-commit 0.3.0's KZG10.

`KZG10.commit(powers, coeffs)` replaces the non-hiding part of `KZG10::commit` (src/kzg10/mod.rs; reached
from /root/reference/benches/marlin.rs:202,311 via MarlinKZG10::commit): leading-zero skip, `into_repr()`
and `VariableBaseMSM::multi_scalar_mul(&powers.powers_of_g[z..], &coeffs)`.  The SRS powers are registered
on the device once (`Powers`); every commit then uploads only the coefficients.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .msm import AffinePoint, RegisteredBases, coord_words, _curve_id


class Powers(RegisteredBases):
    """powers_of_g of a KZG10 committer key, resident in HBM."""

    def __init__(self, curve, powers_of_g, precompute: bool = False):
        if precompute:
            _lib.set_option("msm_precompute", 1)
        try:
            super().__init__(curve, 1, powers_of_g)
        finally:
            if precompute:
                _lib.set_option("msm_precompute", 0)


class KZG10:
    @staticmethod
    def commit(powers: RegisteredBases, coeffs) -> AffinePoint:
        """coeffs: (d + 1, 4) uint64 Montgomery Fr, low degree first (DensePolynomial::coeffs)."""
        c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, _lib.FR_WORDS[powers.curve])
        if len(c) > powers.n:
            raise ValueError("polynomial degree %d exceeds the %d registered powers" % (len(c) - 1, powers.n))
        W = coord_words(powers.curve, 1)
        out = np.zeros(2 * W, dtype=np.uint64)
        oinf = np.zeros(1, dtype=np.uint8)
        _lib.check(_lib.lib().zkm_kzg_commit(powers.handle, ctypes.c_void_p(c.ctypes.data if c.size else 0), len(c),
                                             ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(oinf.ctypes.data)))
        return AffinePoint(powers.curve, 1, out, bool(oinf[0]))

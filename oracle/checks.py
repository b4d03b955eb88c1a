"""TEST INFRASTRUCTURE ONLY (oracle) -- size-independent checkers used by tests/, tools/sweep.py and the
result checks of bench.py.  Never imported by the product path.

  dlog_sum / msm_identity_ok : known-discrete-log identity of an MSM over bases P_i = (a0 + i d) G:
                               sum_i s_i P_i == (sum_i s_i (a0 + i d) mod r) G, right-hand side in exact big-int
                               arithmetic (oracle/py/exact.py)
  horner_ok                  : outputs of an NTT of a short polynomial against Horner evaluation (big ints)
"""
from __future__ import annotations

import numpy as np

from .py import exact
from .py.params import CURVES_BY_ID


def dlog_sum(scal: np.ndarray, a0: int, d: int, r: int, first_index: int = 0) -> int:
    """sum_i s_i (a0 + (first_index + i) d) mod r, exactly and vectorised: scalars are split into 32-bit halves so
    that every numpy accumulation stays below 2^64 (halves < 2^32, indices < 2^27, chunks of 32 products)."""
    scal = np.ascontiguousarray(scal, dtype=np.uint64)
    n, S = scal.shape
    if n == 0:
        return 0
    assert n <= (1 << 27)
    half = scal.view(np.uint32).reshape(n, 2 * S)            # little-endian 32-bit halves
    idx = np.arange(n, dtype=np.uint64)
    tot_s, tot_is = 0, 0
    pad = (-n) % 32
    for j in range(2 * S):
        h = half[:, j].astype(np.uint64)
        tot_s += int(h.sum(dtype=np.uint64)) << (32 * j)
        prod = h * idx
        if pad:
            prod = np.concatenate([prod, np.zeros(pad, dtype=np.uint64)])
        chunks = prod.reshape(-1, 32).sum(axis=1, dtype=np.uint64)
        tot_is += sum(int(v) for v in chunks) << (32 * j)
    return ((a0 + first_index * d) * tot_s + d * tot_is) % r


def msm_identity_ok(curve_id: int, group: int, record: np.ndarray, k: int) -> bool:
    """`record` = the library's result record (2 W coordinate words + infinity flag) == k * G, byte for byte."""
    curve = CURVES_BY_ID[curve_id]
    G = exact.Group(curve, group)
    b, f = exact.point_to_bytes(curve, group, G.mul(G.gen, k % curve.fr.modulus))
    record = np.ascontiguousarray(record, dtype=np.uint64)
    return bool(int(record[-1]) == int(f) and record[:-1].tobytes() == b)


def horner_ok(curve_id: int, log_n: int, coeffs_mont: np.ndarray, evals_at, coset: bool = False) -> bool:
    """`evals_at`: {k: Montgomery limbs of output k} of the size-2^log_n (coset) FFT of the polynomial whose
    Montgomery coefficients are `coeffs_mont` (zero-padded).  Exact big-int Horner at w^k (g w^k for the coset)."""
    from . import capi
    fr = CURVES_BY_ID[curve_id].fr
    coeffs = [fr.from_mont(v) for v in capi.limbs_to_ints(coeffs_mont)]
    d = exact.domain_constants(fr, log_n)
    for k, limbs in evals_at.items():
        pt = pow(d["group_gen"], k, fr.modulus)
        if coset:
            pt = pt * fr.generator % fr.modulus
        got = fr.from_mont(capi.limbs_to_ints(np.asarray(limbs, dtype=np.uint64)[None, :])[0])
        if got != exact.horner_eval(fr, coeffs, pt):
            return False
    return True

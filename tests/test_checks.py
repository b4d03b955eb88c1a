"""The size-independent checkers of oracle/checks.py against plain big-int arithmetic (CPU)."""
import numpy as np

from oracle import capi, checks
from oracle.py import exact
from oracle.py.params import BLS12_381, BW6_761


def test_dlog_sum_matches_big_int():
    rng = np.random.default_rng(1)
    for curve, S in ((BLS12_381, 4), (BW6_761, 6)):
        for n in (0, 1, 31, 32, 33, 1000):
            s = rng.integers(0, 1 << 64, size=(n, S), dtype=np.uint64)
            want = sum(sum(int(s[i, j]) << (64 * j) for j in range(S)) * (7 + (5 + i) * 11) for i in range(n)) % curve.fr.modulus
            assert checks.dlog_sum(s, 7, 11, curve.fr.modulus, first_index=5) == want


def test_msm_identity_checker_accepts_oracle_msm_and_rejects_a_wrong_point():
    n = 200
    a0, d = 0x1234567, 0x89ABCDE
    for cid, g in ((0, 1), (0, 2), (2, 1)):
        bases = capi.progression(cid, g, a0, d, n)
        scal = capi.random_scalars(cid, n, seed=3)
        xy, inf = capi.msm(cid, g, bases, scal)
        rec = np.concatenate([xy, np.array([1 if inf else 0], dtype=np.uint64)])
        k = checks.dlog_sum(scal, a0, d, capi.CURVES[cid].fr.modulus)
        assert checks.msm_identity_ok(cid, g, rec, k)
        assert not checks.msm_identity_ok(cid, g, rec, k + 1)


def test_horner_checker():
    fr = BLS12_381.fr
    x = capi.random_field_elements(0, 1 << 10, seed=4)
    ev = capi.ntt(0, x)
    assert checks.horner_ok(0, 10, x, {k: ev[k] for k in (0, 1, 513, 1023)})
    cev = capi.ntt(0, x, coset=True)
    assert checks.horner_ok(0, 10, x, {k: cev[k] for k in (0, 7)}, coset=True)
    assert not checks.horner_ok(0, 10, x, {3: ev[4]})

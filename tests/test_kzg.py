"""KZG10 commit (hiding term, batched) and open (device witness-polynomial division) -- SURVEY 8f-3 / a14 -- against
the exact restatement in oracle/py/kzg_exact.py and, for the MSM part at larger sizes, the C++ oracle."""
import random

import numpy as np
import pytest

from oracle import capi
from oracle.py import exact, kzg_exact as kx
from oracle.py.params import BLS12_381, BN254, BW6_761

CURVES = [BLS12_381, BN254, BW6_761]


def test_witness_polynomial_is_the_quotient():
    """(X - z) q(X) + p(z) == p(X), coefficient by coefficient (pins the oracle's division)."""
    p = BLS12_381.fr.modulus
    rng = random.Random(1)
    for n in (1, 2, 3, 17, 100):
        c = [rng.randrange(p) for _ in range(n)]
        z = rng.randrange(p)
        q = kx.witness_polynomial(p, c, z)
        assert len(q) == n - 1
        back = [0] * n
        for i, v in enumerate(q):
            back[i + 1] = (back[i + 1] + v) % p
            back[i] = (back[i] - z * v) % p
        back[0] = (back[0] + kx.evaluate(p, c, z)) % p
        assert back == c


@pytest.fixture(scope="module")
def zkm():
    import zkmember_b200 as z
    z.init(0)
    return z


def _mont(fr, vals):
    return capi.ints_to_limbs([fr.to_mont(v) for v in vals], fr.limbs64)


def _pt(curve, ap):
    return exact.point_from_bytes(curve, 1, ap.xy.tobytes(), 1 if ap.infinity else 0)


@pytest.mark.gpu
@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_kzg_commit_hiding_open_small_exact(zkm, curve):
    """Everything against exact big-int arithmetic on a small SRS, incl. zero / constant polynomials and z = 0."""
    from zkmember_b200.kzg import KZG10, Powers
    fr = curve.fr
    p = fr.modulus
    rng = random.Random(7 + curve.curve_id)
    G = exact.Group(curve, 1)
    n = 80
    tau, gamma = rng.randrange(2, p), rng.randrange(2, p)
    powers = [G.mul(G.gen, pow(tau, i, p)) for i in range(n)]
    gpowers = [G.mul(G.gen, gamma * pow(tau, i, p) % p) for i in range(n)]
    arr = lambda pts: np.stack([np.frombuffer(exact.point_to_bytes(curve, 1, Q)[0], dtype=np.uint64) for Q in pts])
    pw, gw = Powers(curve.name, arr(powers)), Powers(curve.name, arr(gpowers), precompute=True)
    try:
        for deg, lead, nb in ((n - 1, 0, 3), (40, 5, 2), (32, 0, 1), (1, 0, 2), (0, 0, 1)):
            c = [rng.randrange(p) for _ in range(deg + 1)]
            for i in range(min(lead, deg + 1)):
                c[i] = 0
            b = [rng.randrange(p) for _ in range(nb)]
            got = KZG10.commit(pw, _mont(fr, c), gw, _mont(fr, b))
            assert _pt(curve, got) == kx.commit(curve, powers, c, gpowers, b)
            for z in (rng.randrange(p), 0, 1):
                pr = KZG10.open(pw, _mont(fr, c), _mont(fr, [z])[0])
                w, _ = kx.open_(curve, powers, c, z)
                assert _pt(curve, pr.w) == w and pr.random_v is None
                prh = KZG10.open(pw, _mont(fr, c), _mont(fr, [z])[0], gw, _mont(fr, b))
                wh, rv = kx.open_(curve, powers, c, z, gpowers, b)
                assert _pt(curve, prh.w) == wh
                assert capi.limbs_to_ints(prh.random_v[None, :])[0] == fr.to_mont(rv)
        # the all-zero polynomial commits to the identity and opens to the identity
        zero = KZG10.commit(pw, _mont(fr, [0, 0, 0]))
        assert zero.infinity
        assert KZG10.open(pw, _mont(fr, [0, 0, 0]), _mont(fr, [5])[0]).w.infinity
    finally:
        pw.release()
        gw.release()


@pytest.mark.gpu
@pytest.mark.parametrize("curve", [BLS12_381, BW6_761], ids=lambda c: c.name)
@pytest.mark.parametrize("n", [1000, 33000])
def test_kzg_open_division_matches_exact_and_commit_matches_oracle(zkm, curve, n):
    """The device division across chunk / thread-run boundaries (n = 33000 spans > 512 x 32 coefficients): the
    opening equals commit(exact quotient) computed by the C++ oracle's MSM."""
    from zkmember_b200.kzg import KZG10, Powers
    fr = curve.fr
    p = fr.modulus
    fid = {0: 1, 1: 3, 2: 5}[curve.curve_id]
    powers = capi.progression(curve.curve_id, 1, 5, 3, n)               # stand-in SRS (any G1 points)
    pw = Powers(curve.name, powers)
    try:
        coeffs_m = capi.random_field_elements(curve.curve_id, n, seed=n)
        coeffs = [fr.from_mont(v) for v in capi.limbs_to_ints(coeffs_m)]
        z = 0x123456789ABCDEF % p
        q = kx.witness_polynomial(p, coeffs, z)
        want_xy, want_inf = capi.msm(curve.curve_id, 1, powers[:n - 1], capi.ints_to_limbs(q, fr.limbs64))
        pr = KZG10.open(pw, coeffs_m, _mont(fr, [z])[0])
        assert pr.w.infinity == want_inf and np.array_equal(pr.w.xy, want_xy)
    finally:
        pw.release()


@pytest.mark.gpu
@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_kzg_commit_batch_equals_single_commits(zkm, curve):
    from zkmember_b200.kzg import KZG10, Powers
    n = 3000
    powers = capi.progression(curve.curve_id, 1, 9, 2, n)
    pw = Powers(curve.name, powers)
    try:
        polys = [capi.random_field_elements(curve.curve_id, m, seed=m) for m in (n, 1777, 1, 64, 2999, 500, 1200, 8, 2048)]
        polys[3][:10] = 0                                        # leading zeros
        polys.append(np.zeros((0, curve.fr.limbs64), dtype=np.uint64))   # empty polynomial -> identity
        got = KZG10.commit_batch(pw, polys)
        assert len(got) == len(polys)
        for g, poly in zip(got, polys):
            one = KZG10.commit(pw, poly)
            assert g == one
        fid_repr = lambda poly: capi.fr_into_repr(curve.curve_id, poly)
        want_xy, want_inf = capi.msm(curve.curve_id, 1, powers[:1777], fid_repr(polys[1]))
        assert np.array_equal(got[1].xy, want_xy) and got[1].infinity == want_inf
        assert got[-1].infinity
    finally:
        pw.release()

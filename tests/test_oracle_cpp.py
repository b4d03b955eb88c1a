"""The C++ restatement of the arkworks algorithms (oracle/cpp) against the exact big-int oracle
(oracle/py/exact.py).  CPU only.  The reference holds no golden vectors for this path
(parity unpinned), so the pins are: exact arithmetic, algorithm-independent
definitions (naive MSM, O(n^2) DFT), known-discrete-log identities and published constants."""
import numpy as np
import pytest

from oracle import capi
from oracle.py import exact
from oracle.py.params import BLS12_381, BN254, BW6_761

CURVES = [BLS12_381, BN254, BW6_761]
FIELDS = {0: BLS12_381.fq, 1: BLS12_381.fr, 2: BN254.fq, 3: BN254.fr, 4: BW6_761.fq, 5: BW6_761.fr}


def _fe(fp, v):
    return capi.ints_to_limbs([fp.to_mont(v)], fp.limbs64)[0]


@pytest.mark.parametrize("fid", [0, 1, 2, 3, 4, 5])
def test_field_ops_match_bigint(fid):
    fp = FIELDS[fid]
    rng = np.random.default_rng(100 + fid)
    p = fp.modulus
    vals = [0, 1, p - 1, 2, (p - 1) // 2] + [int.from_bytes(rng.bytes(104), "little") % p for _ in range(40)]
    for a in vals:
        for b in vals[:8]:
            A, B = _fe(fp, a), _fe(fp, b)
            assert capi.limbs_to_ints(capi.field_op(fid, 0, A, B))[0] == fp.to_mont(a * b % p)
            assert capi.limbs_to_ints(capi.field_op(fid, 1, A, B))[0] == fp.to_mont((a + b) % p)
            assert capi.limbs_to_ints(capi.field_op(fid, 2, A, B))[0] == fp.to_mont((a - b) % p)
        if a:
            assert capi.limbs_to_ints(capi.field_op(fid, 3, A))[0] == fp.to_mont(pow(a, -1, p))
        assert capi.limbs_to_ints(capi.field_op(fid, 4, A))[0] == a


def test_published_montgomery_constants():
    # ark-bls12-381 0.3.0 src/fields/fq.rs: R = 0x15f65ec3fa80e4935c071a97a256ec6d77ce5853705257455f48985753c758baebf4000bc40c0002760900000002fffd
    assert BLS12_381.fq.to_mont(1) == 0x15f65ec3fa80e4935c071a97a256ec6d77ce5853705257455f48985753c758baebf4000bc40c0002760900000002fffd
    assert BLS12_381.fq.inv64 == 0x89f3fffcfffcfffd
    # ark-bls12-381 0.3.0 src/fields/fr.rs: R = 0x1824b159acc5056f998c4fefecbc4ff55884b7fa0003480200000001fffffffe, INV = 0xfffffffeffffffff
    assert BLS12_381.fr.to_mont(1) == 0x1824b159acc5056f998c4fefecbc4ff55884b7fa0003480200000001fffffffe
    assert BLS12_381.fr.inv64 == 0xfffffffeffffffff
    # ark-bn254 0.3.0: Fq INV = 0x87d20782e4866389, Fr INV = 0xc2e1f593efffffff
    assert BN254.fq.inv64 == 0x87d20782e4866389
    assert BN254.fr.inv64 == 0xc2e1f593efffffff
    one = capi.limbs_to_ints(capi.field_op(0, 4, _fe(BLS12_381.fq, 1)))[0]
    assert one == 1


def test_bw6_761_constants():
    """BW6-761 (ark-bw6-761 0.3.0 / ark-bls12-377 0.3.0) is not vendored either: the moduli are pinned by the
    family polynomials, the FFT constants by the published Montgomery limbs of ark-bls12-377's fq.rs."""
    from sympy import isprime
    x = 0x8508c00000000001                                   # the BLS12-377 seed
    r377 = x ** 4 - x ** 2 + 1
    q377 = (x - 1) ** 2 * r377 // 3 + x
    assert BW6_761.fr.modulus == q377 and isprime(q377) and q377.bit_length() == 377
    t0 = x ** 5 - 3 * x ** 4 + 3 * x ** 3 - x + 3            # BW6 family, h_t = 13, h_y = 9
    t, y = t0 + 13 * q377, t0 // 3 + 9 * q377
    assert (t * t + 3 * y * y) % 4 == 0 and (t * t + 3 * y * y) // 4 == BW6_761.fq.modulus
    assert isprime(BW6_761.fq.modulus) and BW6_761.fq.bits == 761
    fr = BW6_761.fr

    def limbs(v, n):
        return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]
    # ark-bls12-377 0.3.0 src/fields/fq.rs: TWO_ADICITY = 46, GENERATOR = -5, R, TWO_ADIC_ROOT_OF_UNITY (Montgomery limbs)
    assert fr.two_adicity == 46 and (fr.modulus - 1) % (1 << 46) == 0 and ((fr.modulus - 1) >> 46) % 2 == 1
    assert fr.generator == fr.modulus - 5
    assert limbs(fr.to_mont(1), 6)[:2] == [0x02cdffffffffff68, 0x51409f837fffffb1]
    assert limbs(fr.to_mont(fr.generator), 6) == [0xfc0b8000000002fa, 0x97d39cf6e000018b, 0x2072420fbfa05044,
                                                   0xcbbcbd50d97c3802, 0x0baf1ec35813f9eb, 0x009974a2c0945ad2]
    assert limbs(fr.to_mont(fr.two_adic_root), 6) == [0x1c104955744e6e0f, 0xf1bd15c3898dd1af, 0x76da78169a7f3950,
                                                       0xee086c1fe367c337, 0xf95564f4cbc1b61f, 0x00f3c1414ef58c54]
    assert pow(fr.two_adic_root, 1 << 45, fr.modulus) == fr.modulus - 1          # order exactly 2^46
    # ark-bw6-761 0.3.0 src/fields/fq.rs: R = 0x0202ffffffff85d5, 0x5a5826358fff8ce7, 0x9e996e43827faade, ...
    assert limbs(BW6_761.fq.to_mont(1), 12)[:3] == [0x0202ffffffff85d5, 0x5a5826358fff8ce7, 0x9e996e43827faade]
    # G1: y^2 = x^3 - 1, G2: y^2 = x^3 + 4, both over Fq; the synthetic generators have order r
    for g in (1, 2):
        G = exact.Group(BW6_761, g)
        assert G.on_curve(G.gen) and G.mul(G.gen, fr.modulus) is None


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 5, 8])
@pytest.mark.parametrize("inverse,coset", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_ntt_matches_definition(curve, log_n, inverse, coset):
    fr = curve.fr
    n = 1 << log_n
    rng = np.random.default_rng(7 * log_n + inverse + 2 * coset)
    x = [int.from_bytes(rng.bytes(56), "little") % fr.modulus for _ in range(n)]
    want = exact.ntt_def(fr, x, bool(inverse), bool(coset)) if n <= 256 else exact.ntt_fast(fr, x, bool(inverse), bool(coset))
    data = capi.ints_to_limbs([fr.to_mont(v) for v in x], fr.limbs64)
    got = capi.ntt(curve.curve_id, data, bool(inverse), bool(coset))
    assert capi.limbs_to_ints(got) == [fr.to_mont(v) for v in want]


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_ntt_large_roundtrip_and_horner(curve):
    fr = curve.fr
    log_n = 14
    n = 1 << log_n
    data = capi.random_field_elements(curve.curve_id, n, seed=0x5EED1000 + log_n)
    ev = capi.ntt(curve.curve_id, data)
    back = capi.ntt(curve.curve_id, ev, inverse=True)
    assert np.array_equal(back, data)
    cev = capi.ntt(curve.curve_id, data, coset=True)
    assert np.array_equal(capi.ntt(curve.curve_id, cev, inverse=True, coset=True), data)
    coeffs = [fr.from_mont(v) for v in capi.limbs_to_ints(data)]
    d = exact.domain_constants(fr, log_n)
    for k in (0, 1, 5, n // 2, n - 1):
        assert fr.from_mont(capi.limbs_to_ints(ev[k:k + 1])[0]) == exact.horner_eval(fr, coeffs, pow(d["group_gen"], k, fr.modulus))
        pt = d["generator"] * pow(d["group_gen"], k, fr.modulus) % fr.modulus
        assert fr.from_mont(capi.limbs_to_ints(cev[k:k + 1])[0]) == exact.horner_eval(fr, coeffs, pt)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_domain_constants(curve):
    fr = curve.fr
    for log_n in (0, 1, 7, 16, fr.two_adicity):
        d = capi.domain(curve.curve_id, log_n)
        e = exact.domain_constants(fr, log_n)
        for k in ("group_gen", "group_gen_inv", "size_inv", "generator", "generator_inv"):
            assert capi.limbs_to_ints(d[k][None, :])[0] == fr.to_mont(e[k]), (log_n, k)
    with pytest.raises(ValueError):
        capi.domain(curve.curve_id, fr.two_adicity + 1)


def _points_to_array(curve, g, pts):
    W = curve.fq.limbs64 * curve.coord_degree(g)
    arr = np.zeros((len(pts), 2 * W), dtype=np.uint64)
    inf = np.zeros(len(pts), dtype=np.uint8)
    for i, P in enumerate(pts):
        b, f = exact.point_to_bytes(curve, g, P)
        arr[i] = np.frombuffer(b, dtype=np.uint64)
        inf[i] = f
    return arr, inf


def _result_point(curve, g, xy, is_inf):
    return exact.point_from_bytes(curve, g, xy.tobytes(), int(is_inf))


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("g", [1, 2])
def test_progression_matches_exact(curve, g):
    G = exact.Group(curve, g)
    want = G.progression(5, 3, 9)
    got = capi.progression(curve.curve_id, g, 5, 3, 9)
    arr, _ = _points_to_array(curve, g, want)
    assert np.array_equal(arr, got)
    assert all(G.on_curve(P) for P in want)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("g", [1, 2])
@pytest.mark.parametrize("n,kind", [(0, "uniform"), (1, "uniform"), (7, "uniform"), (40, "uniform"), (64, "witness"), (33, "small")])
def test_msm_matches_naive(curve, g, n, kind):
    G = exact.Group(curve, g)
    pts = G.progression(11, 7, n)
    # exceptional inputs: points at infinity, duplicates, P and -P
    if n >= 7:
        pts[2] = None
        pts[4] = pts[3]
        pts[6] = G.neg(pts[5])
    scal = capi.random_scalars(curve.curve_id, n, seed=n * 10 + g, kind=kind)
    if n >= 7:
        scal[6] = scal[5]            # s*P + s*(-P) cancels
    arr, inf = _points_to_array(curve, g, pts)
    xy, is_inf = capi.msm(curve.curve_id, g, arr, scal, inf)
    want = G.msm_naive(pts, capi.limbs_to_ints(scal))
    assert _result_point(curve, g, xy, is_inf) == want


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_msm_known_discrete_log_2p12(curve):
    """sum s_i (a0 + i d) G == (sum s_i (a0 + i d) mod r) G -- independent of the MSM algorithm."""
    n = 1 << 12
    a0, d = 0x1234567, 0x89ABCDE
    bases = capi.progression(curve.curve_id, 1, a0, d, n)
    scal = capi.random_scalars(curve.curve_id, n, seed=0x5EED0000 + 12)
    xy, is_inf = capi.msm(curve.curve_id, 1, bases, scal)
    k = sum(s * (a0 + i * d) for i, s in enumerate(capi.limbs_to_ints(scal))) % curve.fr.modulus
    G = exact.Group(curve, 1)
    assert _result_point(curve, 1, xy, is_inf) == G.mul(G.gen, k)


def test_window_bits_examples():
    # SURVEY.md Appendix A.1 examples of c = ln_without_floats(n) + 2
    assert [capi.msm_window_bits(1 << k) for k in (15, 16, 20, 24, 26)] == [12, 13, 15, 18, 19]
    assert capi.msm_window_bits(31) == 3

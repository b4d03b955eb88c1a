"""Groth16 proof assembly and wire format: the host layer's `create_proof` (witness map and five MSMs
on the GPU through the C ABI, `calculate_coeff` / g_c assembly as tiny MSMs, arkworks `CanonicalSerialize`) against
the exact big-int restatement of ark-groth16 0.3.0 src/prover.rs in oracle/py/groth16_exact.py -- proof BYTES equal."""
import random

import numpy as np
import pytest

from oracle import capi
from oracle.py import exact, groth16_exact as gx
from oracle.py.params import BLS12_381, BN254, BW6_761

CURVES = [BLS12_381, BN254, BW6_761]


def _affine(curve, g, P):
    from zkmember_b200.msm import AffinePoint
    b, f = exact.point_to_bytes(curve, g, P)
    return AffinePoint(curve.curve_id, g, np.frombuffer(b, dtype=np.uint64).copy(), bool(f))


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("g", [1, 2])
def test_serialize_matches_exact_restatement(curve, g):
    """Pure host code: compressed points incl. infinity, both y signs, and the Fp2 ordering rule."""
    from zkmember_b200.serialize import serialize_affine
    G = exact.Group(curve, g)
    pts = G.progression(3, 11, 12) + [None]
    pts += [G.neg(P) for P in pts[:6]]
    flags = set()
    for P in pts:
        got = serialize_affine(_affine(curve, g, P))
        want = gx.serialize_affine(curve, g, P)
        assert got == want
        flags.add(got[-1] >> 6)
    assert flags == {0, 1, 2}                                 # NegativeY, Infinity, PositiveY all exercised
    size = (curve.fq.bits + 2 + 7) // 8
    assert len(got) == size + (curve.coord_degree(g) - 1) * ((curve.fq.bits + 7) // 8)


def test_proof_sizes():
    assert len(gx.serialize_proof(BLS12_381, None, None, None)) == 192      # 48 + 96 + 48
    assert len(gx.serialize_proof(BW6_761, None, None, None)) == 288


def _pk(curve, rng, n, num_inputs, m):
    """A structurally valid (not cryptographically meaningful) proving key: arbitrary curve points with a few
    points at infinity and repeats in the b queries, as real proving keys have."""
    G1, G2 = exact.Group(curve, 1), exact.Group(curve, 2)
    p1 = G1.progression(rng.randrange(1, 1 << 20), rng.randrange(1, 1 << 20), 3 * (m + 1) + (n - 1) + (m - num_inputs) + 3)
    p2 = G2.progression(rng.randrange(1, 1 << 20), rng.randrange(1, 1 << 20), (m + 1) + 2)
    it1, it2 = iter(p1), iter(p2)
    pk = {"alpha_g1": next(it1), "beta_g1": next(it1), "delta_g1": next(it1), "beta_g2": next(it2), "delta_g2": next(it2)}
    pk["a_query"] = [next(it1) for _ in range(m + 1)]
    pk["b_g1_query"] = [next(it1) for _ in range(m + 1)]
    pk["b_g2_query"] = [next(it2) for _ in range(m + 1)]
    pk["h_query"] = [next(it1) for _ in range(n - 1)]
    pk["l_query"] = [next(it1) for _ in range(m - num_inputs)]
    for i in (2, 5):
        pk["b_g1_query"][i] = None
        pk["b_g2_query"][i] = None
    pk["b_g1_query"][7] = pk["b_g1_query"][6]
    return pk


@pytest.mark.gpu
@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("precompute", [False, True])
def test_create_proof_bytes_match_exact(curve, precompute):
    import zkmember_b200 as zkm
    from zkmember_b200.groth16 import ProvingKey, create_proof
    zkm.init(0)
    rng = random.Random(1234 + curve.curve_id)
    fr = curve.fr
    log_n, num_inputs, m = 4, 2, 13                          # domain 16, 13 variables besides the constant one
    n = 1 << log_n
    pk = _pk(curve, rng, n, num_inputs, m)
    a, b, c = ([rng.randrange(fr.modulus) for _ in range(n)] for _ in range(3))
    inputs = [rng.randrange(fr.modulus) for _ in range(num_inputs)]
    aux = [rng.choice([0, 1, rng.randrange(fr.modulus)]) for _ in range(m - num_inputs)]
    r, s = rng.randrange(fr.modulus), rng.randrange(fr.modulus)
    A, B, C = gx.create_proof(curve, pk, r, s, a, b, c, inputs, aux)
    want = gx.serialize_proof(curve, A, B, C)

    def arr(g, pts):
        xy = np.stack([np.frombuffer(exact.point_to_bytes(curve, g, P)[0], dtype=np.uint64) for P in pts])
        inf = np.array([exact.point_to_bytes(curve, g, P)[1] for P in pts], dtype=np.uint8)
        return xy, inf
    q = {k: arr(2 if k == "b_g2_query" else 1, pk[k]) for k in ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query")}
    one = lambda g, P: arr(g, [P])[0][0]
    zpk = ProvingKey(curve.name, one(1, pk["alpha_g1"]), one(1, pk["beta_g1"]), one(2, pk["beta_g2"]), one(1, pk["delta_g1"]),
                     one(2, pk["delta_g2"]), q["a_query"][0], q["b_g1_query"][0], q["b_g2_query"][0], q["h_query"][0],
                     q["l_query"][0], infinity={"a": q["a_query"][1], "b_g1": q["b_g1_query"][1], "b_g2": q["b_g2_query"][1],
                                                "h": q["h_query"][1], "l": q["l_query"][1]}, precompute=precompute)
    try:
        mont = lambda v: capi.ints_to_limbs([fr.to_mont(x) for x in v], fr.limbs64)
        canon = lambda v: capi.ints_to_limbs(v, fr.limbs64)
        proof = create_proof(zpk, r, s, mont(a), mont(b), mont(c), canon(inputs), canon(aux))
    finally:
        zpk.release()
    got = proof.serialize()
    assert len(got) == len(want)
    assert got == want

"""Verifier acceptance (the reference's only test across the hot path asserts `Groth16::verify == true`:
/root/reference/src/commitments/pedersen381/mod.rs:64-73, benches/groth16.rs:129).

A genuine Groth16 instance -- R1CS -> QAP -> trusted setup, restated from ark-groth16 0.3.0 in exact big-int
arithmetic (oracle/py/groth16_setup.py) -- is proven (a) by the exact restatement of ark-groth16's prover
(oracle/py/groth16_exact.py: CPU, pins the oracle itself) and (b) by the product: `zkmember_b200.groth16.create_proof`
= witness map (7 NTTs) + five MSMs on the GPU through the C ABI.  Both proofs must satisfy
e(A, B) = e(alpha, beta) e(sum x_j IC_j, gamma) e(C, delta) under the exact pairing of oracle/py/pairing.py, be
rejected for a wrong public input, and be byte-identical to each other."""
import random

import numpy as np
import pytest

from oracle import capi
from oracle.py import exact, groth16_exact as gx, groth16_setup as gs
from oracle.py.params import BLS12_381 as C

P = C.fr.modulus


def _prove_exact(cs, z, par, r, s):
    a, b, c = cs.evaluation_vectors(z, P)
    return gx.create_proof(C, par["pk"], r, s, a, b, c, z[1:cs.num_instance], z[cs.num_instance:])


def test_exact_prover_is_accepted_by_the_pairing_verifier():
    cs, z = gs.cubic_circuit(3, P)
    assert z[1] == 35 and cs.is_satisfied(z, P)
    par = gs.generate_parameters(C, cs, seed=1)
    A, B, Cc = _prove_exact(cs, z, par, random.Random(5).randrange(P), random.Random(6).randrange(P))
    assert gs.verify_proof(C, par["vk"], z[1:cs.num_instance], A, B, Cc)
    assert not gs.verify_proof(C, par["vk"], [36], A, B, Cc)                         # wrong public input
    G1 = exact.Group(C, 1)
    assert not gs.verify_proof(C, par["vk"], z[1:cs.num_instance], A, B, G1.add(Cc, G1.gen))   # tampered proof
    # r = s = 0 (no blinding) is still a valid proof
    A0, B0, C0 = _prove_exact(cs, z, par, 0, 0)
    assert gs.verify_proof(C, par["vk"], z[1:cs.num_instance], A0, B0, C0)


def test_exact_prover_random_gadget_circuit():
    """Boolean witnesses, variables absent from B (points at infinity in the b-queries), empty C rows."""
    cs, z = gs.random_circuit(num_constraints=11, num_inputs=2, seed=7, p=P)
    par = gs.generate_parameters(C, cs, seed=2)
    assert any(q is None for q in par["pk"]["b_g1_query"])
    A, B, Cc = _prove_exact(cs, z, par, 12345, 67890)
    assert gs.verify_proof(C, par["vk"], z[1:cs.num_instance], A, B, Cc)
    bad = list(z[1:cs.num_instance])
    bad[0] = (bad[0] + 1) % P
    assert not gs.verify_proof(C, par["vk"], bad, A, B, Cc)


def _arr(g, pts):
    xy = np.stack([np.frombuffer(exact.point_to_bytes(C, g, Q)[0], dtype=np.uint64) for Q in pts])
    inf = np.array([exact.point_to_bytes(C, g, Q)[1] for Q in pts], dtype=np.uint8)
    return xy, inf


def _gpu_proving_key(pk, precompute):
    from zkmember_b200.groth16 import ProvingKey
    q = {k: _arr(2 if k == "b_g2_query" else 1, pk[k]) for k in ("a_query", "b_g1_query", "b_g2_query", "h_query", "l_query")}
    one = lambda g, Q: _arr(g, [Q])[0][0]
    return ProvingKey(C.name, one(1, pk["alpha_g1"]), one(1, pk["beta_g1"]), one(2, pk["beta_g2"]), one(1, pk["delta_g1"]),
                      one(2, pk["delta_g2"]), q["a_query"][0], q["b_g1_query"][0], q["b_g2_query"][0], q["h_query"][0],
                      q["l_query"][0], infinity={"a": q["a_query"][1], "b_g1": q["b_g1_query"][1], "b_g2": q["b_g2_query"][1],
                                                 "h": q["h_query"][1], "l": q["l_query"][1]}, precompute=precompute)


@pytest.mark.gpu
@pytest.mark.parametrize("circuit,precompute", [("cubic", False), ("cubic", True), ("gadgets", False), ("gadgets", True)])
def test_gpu_proof_is_accepted_by_the_pairing_verifier(circuit, precompute):
    import zkmember_b200 as zkm
    from zkmember_b200.groth16 import create_proof
    zkm.init(0)
    if circuit == "cubic":
        cs, z = gs.cubic_circuit(0xC0FFEE, P)
    else:
        cs, z = gs.random_circuit(num_constraints=27, num_inputs=3, seed=11, p=P)
    par = gs.generate_parameters(C, cs, seed=3)
    rng = random.Random(99)
    r, s = rng.randrange(P), rng.randrange(P)
    a, b, c = cs.evaluation_vectors(z, P)
    inputs, aux = z[1:cs.num_instance], z[cs.num_instance:]
    zpk = _gpu_proving_key(par["pk"], precompute)
    try:
        mont = lambda v: capi.ints_to_limbs([C.fr.to_mont(x) for x in v], C.fr.limbs64)
        canon = lambda v: capi.ints_to_limbs(v, C.fr.limbs64)
        proof = create_proof(zpk, r, s, mont(a), mont(b), mont(c), canon(inputs), canon(aux))
    finally:
        zpk.release()
    pt = lambda ap, g: exact.point_from_bytes(C, g, ap.xy.tobytes(), 1 if ap.infinity else 0)
    A, B, Cc = pt(proof.a, 1), pt(proof.b, 2), pt(proof.c, 1)
    assert gs.verify_proof(C, par["vk"], inputs, A, B, Cc)                            # the verifier accepts
    bad = list(inputs)
    bad[-1] = (bad[-1] + 1) % P
    assert not gs.verify_proof(C, par["vk"], bad, A, B, Cc)
    # and the proof is byte-identical to the exact restatement of ark-groth16's prover on the same (r, s)
    assert proof.serialize() == gx.serialize_proof(C, *_prove_exact(cs, z, par, r, s))

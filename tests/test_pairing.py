"""The exact BLS12-381 pairing of oracle/py/pairing.py (the stand-in for the reference verifier): bilinearity,
non-degeneracy, order r, identity inputs, and the product form a SNARK verifier uses."""
from oracle.py import exact, pairing as pr
from oracle.py.params import BLS12_381


def test_pairing_bilinear_nondegenerate_order_r():
    G1, G2 = exact.Group(BLS12_381, 1), exact.Group(BLS12_381, 2)
    e = pr.pairing(G1.gen, G2.gen)
    assert e != pr.ONE
    assert pr.f12_pow(e, pr.R) == pr.ONE
    a, b = 0x1234567, 0x89ABCDE123
    assert pr.pairing(G1.mul(G1.gen, a), G2.mul(G2.gen, b)) == pr.f12_pow(e, a * b % pr.R)
    assert pr.pairing(None, G2.gen) == pr.ONE and pr.pairing(G1.gen, None) == pr.ONE


def test_pairing_product_form():
    G1, G2 = exact.Group(BLS12_381, 1), exact.Group(BLS12_381, 2)
    a = 987654321
    assert pr.pairing_product_is_one([(G1.mul(G1.gen, a), G2.gen), (G1.neg(G1.gen), G2.mul(G2.gen, a))])
    assert not pr.pairing_product_is_one([(G1.mul(G1.gen, a), G2.gen), (G1.neg(G1.gen), G2.mul(G2.gen, a + 1))])


def test_fq12_field_axioms_spot():
    import random
    rng = random.Random(3)
    x = [rng.randrange(pr.Q) for _ in range(12)]
    y = [rng.randrange(pr.Q) for _ in range(12)]
    assert pr.f12_mul(x, pr.f12_inv(x)) == pr.ONE
    assert pr.f12_mul(pr.f12_mul(x, y), pr.f12_inv(y)) == x
    u = pr.fq2_to_f12((0, 1))                       # u^2 = -1
    assert pr.f12_mul(u, u) == pr.f12([pr.Q - 1])
    w6 = pr.f12_pow(pr.f12([0, 1]), 6)              # w^6 = 1 + u
    assert w6 == pr.fq2_to_f12((1, 1))

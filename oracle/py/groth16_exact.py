"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

Exact big-int restatement of ark-groth16 0.3.0 `create_proof_with_reduction_and_matrices` (src/prover.rs) from
the evaluation vectors onward, and of the compressed `CanonicalSerialize` of the resulting `Proof`
(ark-ec 0.3.0 src/models/short_weierstrass_jacobian.rs `GroupAffine::serialize`, ark-ff 0.3.0
`serialize_with_flags`, ark-serialize 0.3.0 `SWFlags`).  Reached in the reference from
/root/reference/benches/groth16.rs:115 (prove) and /root/reference/src/main.rs:164-169 (proof bytes on the wire).

PARITY UNPINNED against arkworks binaries (no Rust toolchain, no golden proof in the reference: its proofs are
randomised by `rng`); what this pins is the host-side assembly and byte layout of the product against an
independent statement of the same published algorithms.
"""
from typing import List, Sequence

from . import exact
from .params import CurveParams


def witness_map(curve: CurveParams, a: Sequence[int], b: Sequence[int], c: Sequence[int]) -> List[int]:
    """R1CStoQAP::witness_map after the matrix-vector products (src/r1cs_to_qap.rs): canonical integers in and out."""
    fr = curve.fr
    p = fr.modulus
    n = len(a)
    log_n = n.bit_length() - 1
    a = exact.ntt_def(fr, exact.ntt_def(fr, a, inverse=True), coset=True)
    b = exact.ntt_def(fr, exact.ntt_def(fr, b, inverse=True), coset=True)
    c = exact.ntt_def(fr, exact.ntt_def(fr, c, inverse=True), coset=True)
    zinv = pow(pow(fr.generator, n, p) - 1, -1, p)          # 1 / Z_H(g),  Z_H(X) = X^n - 1
    ab = [((x * y - z) * zinv) % p for x, y, z in zip(a, b, c)]
    assert 1 << log_n == n
    return exact.ntt_def(fr, ab, inverse=True, coset=True)


def calculate_coeff(G: exact.Group, initial, query: Sequence, vk_param, assignment: Sequence[int]):
    """src/prover.rs calculate_coeff: el = query[0]; acc = msm(query[1..], assignment);
    res = initial; res += el; res += acc; res += vk_param."""
    acc = G.msm_naive(list(query[1:]), list(assignment))
    res = G.add(initial, query[0])
    res = G.add(res, acc)
    return G.add(res, vk_param)


def create_proof(curve: CurveParams, pk: dict, r: int, s: int, a, b, c, input_assignment, aux_assignment):
    """pk: dict of affine points / lists (alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, a_query, b_g1_query,
    b_g2_query, h_query, l_query).  Returns (A, B, C) as exact affine points."""
    G1, G2 = exact.Group(curve, 1), exact.Group(curve, 2)
    h = witness_map(curve, a, b, c)
    h_acc = G1.msm_naive(list(pk["h_query"]), h[:len(pk["h_query"])])
    l_aux_acc = G1.msm_naive(list(pk["l_query"]), list(aux_assignment))
    r_s_delta_g1 = G1.mul(pk["delta_g1"], (r * s) % curve.fr.modulus)
    assignment = list(input_assignment) + list(aux_assignment)
    g_a = calculate_coeff(G1, G1.mul(pk["delta_g1"], r), pk["a_query"], pk["alpha_g1"], assignment)
    g1_b = calculate_coeff(G1, G1.mul(pk["delta_g1"], s), pk["b_g1_query"], pk["beta_g1"], assignment)
    g2_b = calculate_coeff(G2, G2.mul(pk["delta_g2"], s), pk["b_g2_query"], pk["beta_g2"], assignment)
    g_c = G1.mul(g_a, s)
    g_c = G1.add(g_c, G1.mul(g1_b, r))
    g_c = G1.add(g_c, G1.neg(r_s_delta_g1))
    g_c = G1.add(g_c, l_aux_acc)
    g_c = G1.add(g_c, h_acc)
    return g_a, g2_b, g_c


def _field_bytes_with_flags(q_bits: int, value: int, flag_bits: int, mask: int) -> bytes:
    size = (q_bits + flag_bits + 7) // 8                     # buffer_byte_size(MODULUS_BITS + F::BIT_SIZE)
    out = bytearray(value.to_bytes(size, "little"))
    out[size - 1] |= mask
    return bytes(out)


def serialize_affine(curve: CurveParams, g: int, P) -> bytes:
    """GroupAffine::serialize: infinity -> zero x with SWFlags::Infinity (1 << 6); else x with
    SWFlags::from_y_sign(y > -y) (PositiveY = 1 << 7, NegativeY = 0)."""
    q = curve.fq.modulus
    bits = curve.fq.bits
    deg = curve.coord_degree(g)
    if P is None:
        x = 0 if deg == 1 else (0, 0)
        mask = 1 << 6
    else:
        x, y = P
        if deg == 1:
            positive = y > (q - y) % q
        else:                                                # QuadExtField::cmp: c1 first, then c0
            ny = ((q - y[0]) % q, (q - y[1]) % q)
            positive = (y[1], y[0]) > (ny[1], ny[0])
        mask = (1 << 7) if positive else 0
    if deg == 1:
        return _field_bytes_with_flags(bits, x, 2, mask)
    return _field_bytes_with_flags(bits, x[0], 0, 0) + _field_bytes_with_flags(bits, x[1], 2, mask)


def serialize_proof(curve: CurveParams, A, B, C) -> bytes:
    return serialize_affine(curve, 1, A) + serialize_affine(curve, 2, B) + serialize_affine(curve, 1, C)

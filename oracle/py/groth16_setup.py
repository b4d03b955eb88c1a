"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

A small but GENUINE Groth16 instance in exact big-int arithmetic, so that proofs assembled from the GPU's MSM / NTT
outputs can be checked against the pairing equation of the verifier (the reference's only test across the hot path
asserts exactly `verify == true`: /root/reference/src/commitments/pedersen381/mod.rs:64-73).

Restated from ark-groth16 0.3.0 (pin /root/reference/Cargo.lock:286-287):
  * `R1CStoQAP::instance_map_with_evaluation` (src/r1cs_to_qap.rs): with u_i = L_i(t) the Lagrange coefficients of the
    domain H (|H| = next_pow2(num_constraints + num_instance)), a_j(t) = sum_i u_i A[i][j] (+ u_{nc + j} for instance
    variable j: the input-consistency rows), b_j(t) = sum_i u_i B[i][j], c_j(t) = sum_i u_i C[i][j], Z(t) = t^n - 1;
  * `generate_parameters` (src/generator.rs): a_query = a_j(t) G1, b_g1/b_g2_query = b_j(t) G1/G2,
    h_query[i] = (Z(t) / delta) t^i G1 for i < n - 1, l_query = ((beta a_j + alpha b_j + c_j) / delta) G1 for witness
    variables, gamma_abc_g1 = ((beta a_j + alpha b_j + c_j) / gamma) G1 for instance variables;
  * `verify_proof` (src/verifier.rs): e(A, B) = e(alpha_g1, beta_g2) e(sum_j x_j gamma_abc_g1[j], gamma_g2) e(C, delta_g2);
  * the evaluation vectors `witness_map` starts from (src/r1cs_to_qap.rs): a[i] = <A_i, z>, b[i] = <B_i, z>,
    c[i] = <C_i, z> for i < nc and a[nc + j] = z_j for the instance variables.
The toxic waste (t, alpha, beta, gamma, delta) comes from a seeded PRNG -- this is a test fixture, not a ceremony.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

from . import exact
from .pairing import pairing_product_is_one
from .params import BLS12_381, CurveParams

Row = Dict[int, int]            # sparse linear combination: variable index -> coefficient


@dataclass
class R1CS:
    """Variables are ordered [1, public inputs..., witness...] (ark-relations: instance first, then witness)."""
    num_instance: int            # including the constant one
    num_witness: int
    A: List[Row]
    B: List[Row]
    C: List[Row]

    @property
    def num_constraints(self) -> int:
        return len(self.A)

    @property
    def num_variables(self) -> int:
        return self.num_instance + self.num_witness

    def domain_log(self) -> int:
        n = self.num_constraints + self.num_instance
        return max((n - 1).bit_length(), 1)

    def is_satisfied(self, z: Sequence[int], p: int) -> bool:
        dot = lambda row: sum(c * z[j] for j, c in row.items()) % p
        return all(dot(a) * dot(b) % p == dot(c) for a, b, c in zip(self.A, self.B, self.C))

    def evaluation_vectors(self, z: Sequence[int], p: int) -> Tuple[List[int], List[int], List[int]]:
        n = 1 << self.domain_log()
        dot = lambda row: sum(c * z[j] for j, c in row.items()) % p
        a, b, c = [0] * n, [0] * n, [0] * n
        for i in range(self.num_constraints):
            a[i], b[i], c[i] = dot(self.A[i]), dot(self.B[i]), dot(self.C[i])
        for j in range(self.num_instance):
            a[self.num_constraints + j] = z[j] % p
        return a, b, c


def cubic_circuit(x: int, p: int) -> Tuple[R1CS, List[int]]:
    """x^3 + x + 5 = out: three constraints, instance (1, out), witness (x, x^2, x^3)."""
    v1, v2 = x * x % p, x * x * x % p
    out = (v2 + x + 5) % p
    cs = R1CS(num_instance=2, num_witness=3,
              A=[{2: 1}, {3: 1}, {4: 1, 2: 1, 0: 5}],
              B=[{2: 1}, {2: 1}, {0: 1}],
              C=[{3: 1}, {4: 1}, {1: 1}])
    return cs, [1, out, x % p, v1, v2]


def random_circuit(num_constraints: int, num_inputs: int, seed: int, p: int, zero_one: float = 0.5) -> Tuple[R1CS, List[int]]:
    """A satisfiable random R1CS shaped like a gadget circuit: constraint i multiplies two sparse combinations of
    earlier variables and defines a new witness variable; a fraction of the witness is boolean (0/1 scalars, as in
    zkMember's bit decompositions), some variables never occur in B (points at infinity in the b-queries)."""
    rng = random.Random(seed)
    z = [1] + [rng.randrange(p) for _ in range(num_inputs)]
    A, B, C = [], [], []
    ni = 1 + num_inputs
    for i in range(num_constraints):
        nv = len(z)
        if rng.random() < zero_one:
            # booleanity: b * (1 - b) = 0 for a fresh boolean witness b
            bit = rng.randrange(2)
            z.append(bit)
            A.append({nv: 1})
            B.append({0: 1, nv: p - 1})
            C.append({})
        else:
            ra = {rng.randrange(nv): rng.randrange(1, p) for _ in range(rng.randrange(1, 4))}
            rb = {rng.randrange(min(nv, ni + 3)): rng.randrange(1, p) for _ in range(rng.randrange(1, 3))}
            dot = lambda row: sum(c * z[j] for j, c in row.items()) % p
            z.append(dot(ra) * dot(rb) % p)
            A.append(ra)
            B.append(rb)
            C.append({nv: 1})
    cs = R1CS(num_instance=ni, num_witness=len(z) - ni, A=A, B=B, C=C)
    assert cs.is_satisfied(z, p)
    return cs, z


def lagrange_at(fr, log_n: int, t: int) -> List[int]:
    """u_i = L_i(t) over H = <w>: Z(t) / (n (t - w^i)) * w^i   (evaluate_all_lagrange_coefficients, t outside H)."""
    p = fr.modulus
    n = 1 << log_n
    w = exact.domain_constants(fr, log_n)["group_gen"]
    zt = (pow(t, n, p) - 1) % p
    ninv = pow(n, -1, p)
    out, wi = [], 1
    for _ in range(n):
        out.append(zt * ninv % p * wi % p * pow((t - wi) % p, -1, p) % p)
        wi = wi * w % p
    return out


def generate_parameters(curve: CurveParams, cs: R1CS, seed: int) -> dict:
    """Proving key + verifying key as exact affine points (None = point at infinity)."""
    fr = curve.fr
    p = fr.modulus
    rng = random.Random(seed)
    G1, G2 = exact.Group(curve, 1), exact.Group(curve, 2)
    log_n = cs.domain_log()
    n = 1 << log_n
    t = rng.randrange(2, p)
    while pow(t, n, p) == 1:
        t = rng.randrange(2, p)
    alpha, beta, gamma, delta = (rng.randrange(1, p) for _ in range(4))
    u = lagrange_at(fr, log_n, t)
    nv = cs.num_variables
    a, b, c = [0] * nv, [0] * nv, [0] * nv
    for j in range(cs.num_instance):
        a[j] = u[cs.num_constraints + j]
    for i in range(cs.num_constraints):
        for j, coef in cs.A[i].items():
            a[j] = (a[j] + u[i] * coef) % p
        for j, coef in cs.B[i].items():
            b[j] = (b[j] + u[i] * coef) % p
        for j, coef in cs.C[i].items():
            c[j] = (c[j] + u[i] * coef) % p
    zt = (pow(t, n, p) - 1) % p
    dinv, ginv = pow(delta, -1, p), pow(gamma, -1, p)
    mul1 = lambda k: G1.mul(G1.gen, k % p) if k % p else None
    mul2 = lambda k: G2.mul(G2.gen, k % p) if k % p else None
    abc = [(beta * a[j] + alpha * b[j] + c[j]) % p for j in range(nv)]
    pk = {
        "alpha_g1": mul1(alpha), "beta_g1": mul1(beta), "beta_g2": mul2(beta), "delta_g1": mul1(delta), "delta_g2": mul2(delta),
        "a_query": [mul1(v) for v in a], "b_g1_query": [mul1(v) for v in b], "b_g2_query": [mul2(v) for v in b],
        "h_query": [mul1(zt * dinv % p * pow(t, i, p)) for i in range(n - 1)],
        "l_query": [mul1(abc[j] * dinv) for j in range(cs.num_instance, nv)],
    }
    vk = {"alpha_g1": pk["alpha_g1"], "beta_g2": pk["beta_g2"], "gamma_g2": mul2(gamma), "delta_g2": pk["delta_g2"],
          "gamma_abc_g1": [mul1(abc[j] * ginv) for j in range(cs.num_instance)]}
    return {"pk": pk, "vk": vk, "log_n": log_n}


def verify_proof(curve: CurveParams, vk: dict, public_inputs: Sequence[int], A, B, C) -> bool:
    """ark_groth16::verify_proof: e(A, B) == e(alpha, beta) e(sum x_j gamma_abc_j, gamma) e(C, delta), as one product."""
    assert curve is BLS12_381, "the exact pairing is written for BLS12-381"
    G1 = exact.Group(curve, 1)
    acc = vk["gamma_abc_g1"][0]
    for x, P in zip(public_inputs, vk["gamma_abc_g1"][1:]):
        acc = G1.add(acc, G1.mul(P, x % curve.fr.modulus) if P is not None else None)
    if A is None or B is None:
        return False
    return pairing_product_is_one([(A, B), (G1.neg(vk["alpha_g1"]), vk["beta_g2"]), (G1.neg(acc), vk["gamma_g2"]),
                                   (G1.neg(C) if C is not None else None, vk["delta_g2"])])

"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

Exact big-int restatement of the polynomial side of ark-poly-commit 0.3.0's KZG10 (src/kzg10/mod.rs; pin
/root/reference/Cargo.lock:352-353; reached from /root/reference/benches/marlin.rs:202,311 via MarlinKZG10):

  commit(p)            = sum_i p_i * powers_of_g[i]                       (skip_leading_zeros is only an optimisation)
                         [+ sum_i b_i * powers_of_gamma_g[i]  for a blinding polynomial b  -- the hiding term]
  compute_witness_polynomial(p, z) = p / (X - z)   (DensePolynomial division, remainder p(z) dropped)
  open(p, z)           = Proof { w: commit(witness) [+ hiding witness over powers_of_gamma_g], random_v: b(z) }

Coefficients are canonical integers here; group arithmetic is oracle/py/exact.py's affine big-int group law.
"""
from typing import List, Optional, Sequence, Tuple

from . import exact
from .params import CurveParams


def evaluate(p: int, coeffs: Sequence[int], z: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * z + c) % p
    return acc


def witness_polynomial(p: int, coeffs: Sequence[int], z: int) -> List[int]:
    """Quotient of coeffs(X) by (X - z): q_{i-1} = c_i + z q_i (synthetic division); len(coeffs) - 1 coefficients."""
    n = len(coeffs)
    if n <= 1:
        return []
    q = [0] * (n - 1)
    acc = 0
    for i in range(n - 1, 0, -1):
        acc = (coeffs[i] + z * acc) % p
        q[i - 1] = acc
    return q


def commit(curve: CurveParams, powers: Sequence, coeffs: Sequence[int], gamma_powers: Optional[Sequence] = None,
           blinding: Optional[Sequence[int]] = None):
    G = exact.Group(curve, 1)
    c = G.msm_naive(list(powers[:len(coeffs)]), list(coeffs))
    if gamma_powers is not None and blinding is not None:
        c = G.add(c, G.msm_naive(list(gamma_powers[:len(blinding)]), list(blinding)))
    return c


def open_(curve: CurveParams, powers: Sequence, coeffs: Sequence[int], z: int, gamma_powers: Optional[Sequence] = None,
          blinding: Optional[Sequence[int]] = None) -> Tuple[object, Optional[int]]:
    p = curve.fr.modulus
    G = exact.Group(curve, 1)
    wq = witness_polynomial(p, coeffs, z)
    w = G.msm_naive(list(powers[:len(wq)]), wq)
    random_v = None
    if gamma_powers is not None and blinding is not None:
        hq = witness_polynomial(p, blinding, z)
        w = G.add(w, G.msm_naive(list(gamma_powers[:len(hq)]), hq))
        random_v = evaluate(p, blinding, z)
    return w, random_v

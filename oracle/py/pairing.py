"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

Exact big-int ate pairing on BLS12-381, written from the definition (no tower tricks, no sparse lines) so that
it is easy to audit: it stands in for the reference's ONLY test that crosses the hot path, which asserts that the
verifier accepts (/root/reference/src/commitments/pedersen381/mod.rs:64-73 `Groth16::verify(..) == true`,
/root/reference/benches/groth16.rs:129).  arkworks itself cannot run here (no Rust toolchain), so acceptance is
checked against the pairing EQUATION of Groth16 with this pairing.

Construction
  Fq12 = Fq[w] / (w^12 - 2 w^6 + 2).  With xi = 1 + u (u^2 = -1) one has w^6 = xi, i.e. u = w^6 - 1, so
  Fq2 = Fq[u]/(u^2 + 1) embeds as c0 + c1 u -> (c0 - c1) + c1 w^6.
  G2 lives on the M-type sextic twist E': y^2 = x^3 + 4 xi; the untwist E'(Fq2) -> E(Fq12) is
  (x', y') -> (x' / w^2, y' / w^3)   [(y'/w^3)^2 = (x'^3 + 4 xi)/xi = (x'/w^2)^3 + 4].
  e(P, Q) = f_{|z|, Q}(P)^((q^12 - 1)/r) inverted (z = -0xd201000000010000 is negative), Miller loop in affine
  coordinates over Fq12, vertical lines dropped (the x-coordinates of untwisted points lie in Fq6, which the final
  exponentiation kills).  Any fixed power of the Tate pairing is bilinear and non-degenerate, which is all the
  Groth16 equation needs; tests/test_pairing.py checks bilinearity, non-degeneracy and order r.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

from .params import BLS12_381

Q = BLS12_381.fq.modulus
R = BLS12_381.fr.modulus
Z_ABS = 0xd201000000010000          # |z|; the BLS parameter is -|z|
DEG = 12
# w^12 = 2 w^6 - 2
FINAL_EXP = (Q ** 12 - 1) // R

Fq12 = List[int]                     # 12 coefficients, little-endian in w


def f12(c: Sequence[int]) -> Fq12:
    return [int(v) % Q for v in c] + [0] * (DEG - len(c))


ONE = f12([1])
ZERO = f12([0])


def f12_add(a, b):
    return [(x + y) % Q for x, y in zip(a, b)]


def f12_sub(a, b):
    return [(x - y) % Q for x, y in zip(a, b)]


def f12_scale(a, k: int):
    return [x * k % Q for x in a]


def f12_mul(a, b):
    t = [0] * (2 * DEG - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    # reduce: w^(12+k) = 2 w^(6+k) - 2 w^k
    for k in range(2 * DEG - 2, DEG - 1, -1):
        v = t[k]
        if v:
            t[k - 6] += 2 * v
            t[k - 12] -= 2 * v
    return [v % Q for v in t[:DEG]]


def f12_sqr(a):
    return f12_mul(a, a)


def _poly_trim(p):
    while p and p[-1] == 0:
        p.pop()
    return p


def _poly_divmod(a, b):
    a = a[:]
    out = [0] * max(len(a) - len(b) + 1, 1)
    inv_lead = pow(b[-1], -1, Q)
    for k in range(len(a) - len(b), -1, -1):
        c = a[k + len(b) - 1] * inv_lead % Q
        out[k] = c
        if c:
            for j, y in enumerate(b):
                a[k + j] = (a[k + j] - c * y) % Q
    return _poly_trim(out), _poly_trim(a[:len(b) - 1])


def f12_inv(a):
    """Extended Euclid on polynomials over Fq modulo w^12 - 2 w^6 + 2 (irreducible, so every a != 0 is a unit)."""
    mod = [2] + [0] * 5 + [Q - 2] + [0] * 5 + [1]
    r0, r1 = mod, _poly_trim(list(a))
    if not r1:
        raise ZeroDivisionError("Fq12 inverse of zero")
    s0, s1 = [], [1]
    while len(r1) > 1:
        qt, rem = _poly_divmod(r0, r1)
        # s0 - qt * s1
        prod = [0] * (len(qt) + len(s1) - 1) if qt and s1 else []
        for i, x in enumerate(qt):
            for j, y in enumerate(s1):
                prod[i + j] = (prod[i + j] + x * y) % Q
        ns = [0] * max(len(s0), len(prod))
        for i in range(len(ns)):
            ns[i] = ((s0[i] if i < len(s0) else 0) - (prod[i] if i < len(prod) else 0)) % Q
        r0, r1, s0, s1 = r1, rem, s1, _poly_trim(ns)
        if not r1:
            raise ZeroDivisionError("Fq12 element is not invertible")
    c = pow(r1[0], -1, Q)
    res = [x * c % Q for x in s1] + [0] * DEG
    # s1 may have degree up to 11; fold anything above (cannot happen, kept for safety)
    return f12_mul(res[:DEG], ONE)


def f12_pow(a, e: int):
    res = ONE
    base = a
    while e:
        if e & 1:
            res = f12_mul(res, base)
        base = f12_mul(base, base)
        e >>= 1
    return res


def fq2_to_f12(c) -> Fq12:
    c0, c1 = c
    out = [0] * DEG
    out[0] = (c0 - c1) % Q
    out[6] = c1 % Q
    return out


_W = f12([0, 1])
_W2_INV = f12_inv(f12_mul(_W, _W))
_W3_INV = f12_inv(f12_mul(f12_mul(_W, _W), _W))


def untwist(Qp) -> Tuple[Fq12, Fq12]:
    """E'(Fq2) -> E(Fq12): (x', y') -> (x'/w^2, y'/w^3)."""
    x, y = Qp
    return f12_mul(fq2_to_f12(x), _W2_INV), f12_mul(fq2_to_f12(y), _W3_INV)


def _line(T, S, P):
    """Value at P of the line through T and S (tangent when T == S); returns (value, T + S).  All in E(Fq12)."""
    (x1, y1), (x2, y2) = T, S
    xp, yp = P
    if x1 != x2:
        lam = f12_mul(f12_sub(y2, y1), f12_inv(f12_sub(x2, x1)))
    elif y1 == y2:
        lam = f12_mul(f12_scale(f12_sqr(x1), 3), f12_inv(f12_scale(y1, 2)))
    else:
        return f12_sub(xp, x1), None           # vertical line: T + S = O
    x3 = f12_sub(f12_sub(f12_sqr(lam), x1), x2)
    y3 = f12_sub(f12_mul(lam, f12_sub(x1, x3)), y1)
    val = f12_sub(f12_sub(yp, y1), f12_mul(lam, f12_sub(xp, x1)))
    return val, (x3, y3)


def miller_loop(P, Qp) -> Fq12:
    """f_{|z|, psi(Q)}(P) for P in E(Fq) (affine ints) and Q in E'(Fq2) (affine pairs); identity inputs give 1."""
    if P is None or Qp is None:
        return ONE
    Pe = (f12([P[0]]), f12([P[1]]))
    Qe = untwist(Qp)
    T = Qe
    f = ONE
    for bit in bin(Z_ABS)[3:]:
        val, T = _line(T, T, Pe)
        f = f12_mul(f12_sqr(f), val)
        if bit == "1":
            val, T = _line(T, Qe, Pe)
            f = f12_mul(f, val)
    return f


def final_exponentiation(f: Fq12) -> Fq12:
    return f12_pow(f, FINAL_EXP)


def pairing(P, Qp) -> Fq12:
    """e(P, Q); the loop runs over |z| and z < 0, hence the inversion."""
    return final_exponentiation(f12_inv(miller_loop(P, Qp)))


def pairing_product_is_one(pairs) -> bool:
    """prod e(P_i, Q_i) == 1 with ONE final exponentiation (the shape of a SNARK verifier)."""
    f = ONE
    for P, Qp in pairs:
        f = f12_mul(f, miller_loop(P, Qp))
    return final_exponentiation(f) == ONE

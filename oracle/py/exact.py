"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

Exact big-integer ground truth for the hot path, independent of any
particular algorithm:

  * `msm_naive`   : sum_i s_i * P_i by affine double-and-add; the mathematical
                    definition of ark-ec 0.3.0 `VariableBaseMSM::multi_scalar_mul`
                    (src/msm/variable_base.rs; reached from
                    /root/reference/benches/groth16.rs:115) normalised with
                    `into_affine()`.
  * `ntt_def`     : X[k] = sum_j x_j w^(jk), the contract of ark-poly 0.3.0
                    `Radix2EvaluationDomain::fft_in_place`
                    (src/domain/radix2/fft.rs) and its
                    ifft / coset_fft / coset_ifft variants.
  * `ntt_fast`    : the same map by a recursive radix-2 split (for 2^10..2^16).

Byte formats are exactly arkworks': field elements are value*R mod p in
little-endian u64 limbs (R = 2^(64*limbs)), scalars are canonical little-endian
(`into_repr()`), affine points are (x, y) plus a separate infinity flag.
PARITY UNPINNED against arkworks binaries -- see params.py.
"""
from __future__ import annotations

import random
from typing import Iterable, List, Optional, Sequence, Tuple

from .params import CurveParams, FieldParams

# ------------------------------------------------------------------ Fq / Fq2 helpers
# A "field context" is (q, deg) where deg 1 -> ints, deg 2 -> tuples (c0, c1), u^2 = -1.


class Fq2Ops:
    def __init__(self, q: int):
        self.q = q
        self.zero = (0, 0)
        self.one = (1, 0)

    def add(self, a, b):
        return ((a[0] + b[0]) % self.q, (a[1] + b[1]) % self.q)

    def sub(self, a, b):
        return ((a[0] - b[0]) % self.q, (a[1] - b[1]) % self.q)

    def neg(self, a):
        return ((-a[0]) % self.q, (-a[1]) % self.q)

    def mul(self, a, b):
        q = self.q
        return ((a[0] * b[0] - a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)

    def inv(self, a):
        q = self.q
        n = pow((a[0] * a[0] + a[1] * a[1]) % q, -1, q)
        return (a[0] * n % q, (-a[1] * n) % q)

    def from_int(self, k: int):
        return (k % self.q, 0)

    def is_zero(self, a):
        return a[0] == 0 and a[1] == 0


class FqOps:
    def __init__(self, q: int):
        self.q = q
        self.zero = 0
        self.one = 1

    def add(self, a, b):
        return (a + b) % self.q

    def sub(self, a, b):
        return (a - b) % self.q

    def neg(self, a):
        return (-a) % self.q

    def mul(self, a, b):
        return (a * b) % self.q

    def inv(self, a):
        return pow(a, -1, self.q)

    def from_int(self, k: int):
        return k % self.q

    def is_zero(self, a):
        return a == 0


# ------------------------------------------------------------------ affine short-Weierstrass (a = 0)
# A point is None (infinity) or (x, y).


class Group:
    """y^2 = x^3 + b over Fq (G1) or Fq2 (G2), a = 0."""

    def __init__(self, curve: CurveParams, g: int):
        self.curve = curve
        self.g = g
        if g == 1:
            self.F = FqOps(curve.fq.modulus)
            self.b = curve.b_g1
            self.gen = curve.g1
        elif curve.g2_over_fq:       # BW6-761: G2 is a curve over Fq as well
            self.F = FqOps(curve.fq.modulus)
            self.b = curve.b_g2
            self.gen = curve.g2
        else:
            self.F = Fq2Ops(curve.fq.modulus)
            self.b = curve.b_g2
            self.gen = curve.g2

    def on_curve(self, P) -> bool:
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.mul(y, y) == F.add(F.mul(F.mul(x, x), x), self.b)

    def neg(self, P):
        if P is None:
            return None
        return (P[0], self.F.neg(P[1]))

    def add(self, P, Q):
        F = self.F
        if P is None:
            return Q
        if Q is None:
            return P
        x1, y1 = P
        x2, y2 = Q
        if x1 == x2:
            if y1 == y2 and not F.is_zero(y1):
                return self.double(P)
            return None
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
        x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    def double(self, P):
        F = self.F
        if P is None:
            return None
        x1, y1 = P
        if F.is_zero(y1):
            return None
        three = F.from_int(3)
        two = F.from_int(2)
        lam = F.mul(F.mul(three, F.mul(x1, x1)), F.inv(F.mul(two, y1)))
        x3 = F.sub(F.mul(lam, lam), F.mul(two, x1))
        y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
        return (x3, y3)

    def mul(self, P, k: int):
        if k < 0:
            return self.mul(self.neg(P), -k)
        R = None
        A = P
        while k:
            if k & 1:
                R = self.add(R, A)
            A = self.double(A)
            k >>= 1
        return R

    def msm_naive(self, bases: Sequence, scalars: Sequence[int]):
        n = min(len(bases), len(scalars))
        acc = None
        for i in range(n):
            if scalars[i] and bases[i] is not None:
                acc = self.add(acc, self.mul(bases[i], scalars[i]))
        return acc

    # arithmetic progression of known discrete logs: P_i = (a0 + i*d) * G
    def progression(self, a0: int, d: int, n: int) -> List:
        out = []
        P = self.mul(self.gen, a0)
        D = self.mul(self.gen, d)
        for _ in range(n):
            out.append(P)
            P = self.add(P, D)
        return out


# ------------------------------------------------------------------ byte formats
def fe_to_bytes(fp: FieldParams, x: int) -> bytes:
    """Montgomery, little-endian u64 limbs (== little-endian bytes)."""
    return fp.to_mont(x).to_bytes(8 * fp.limbs64, "little")


def fe_from_bytes(fp: FieldParams, b: bytes) -> int:
    return fp.from_mont(int.from_bytes(b, "little"))


def scalar_to_bytes(fp: FieldParams, s: int) -> bytes:
    return int(s).to_bytes(8 * fp.limbs64, "little")


def coord_to_bytes(curve: CurveParams, g: int, c) -> bytes:
    if curve.coord_degree(g) == 1:
        return fe_to_bytes(curve.fq, c)
    return fe_to_bytes(curve.fq, c[0]) + fe_to_bytes(curve.fq, c[1])


def coord_from_bytes(curve: CurveParams, g: int, b: bytes):
    w = 8 * curve.fq.limbs64
    if curve.coord_degree(g) == 1:
        return fe_from_bytes(curve.fq, b[:w])
    return (fe_from_bytes(curve.fq, b[:w]), fe_from_bytes(curve.fq, b[w:2 * w]))


def point_to_bytes(curve: CurveParams, g: int, P) -> Tuple[bytes, int]:
    """(xy bytes, infinity flag).  Infinity is encoded like ark-ec's
    GroupAffine::zero(): x = 0, y = 1 (Montgomery one), flag = 1."""
    if P is None:
        if curve.coord_degree(g) == 1:
            return coord_to_bytes(curve, g, 0) + coord_to_bytes(curve, g, 1), 1
        return coord_to_bytes(curve, g, (0, 0)) + coord_to_bytes(curve, g, (1, 0)), 1
    return coord_to_bytes(curve, g, P[0]) + coord_to_bytes(curve, g, P[1]), 0


def point_from_bytes(curve: CurveParams, g: int, b: bytes, inf: int):
    if inf:
        return None
    w = 8 * curve.fq.limbs64 * curve.coord_degree(g)
    return (coord_from_bytes(curve, g, b[:w]), coord_from_bytes(curve, g, b[w:2 * w]))


# ------------------------------------------------------------------ NTT
def domain_constants(fr: FieldParams, log_n: int):
    """Radix2EvaluationDomain::new (ark-poly 0.3.0 src/domain/radix2/mod.rs):
    group_gen = TWO_ADIC_ROOT^(2^(TWO_ADICITY - log_n))."""
    if log_n > fr.two_adicity:
        raise ValueError("log_n exceeds two-adicity")
    n = 1 << log_n
    w = pow(fr.two_adic_root, 1 << (fr.two_adicity - log_n), fr.modulus)
    return dict(
        size=n,
        group_gen=w,
        group_gen_inv=pow(w, -1, fr.modulus),
        size_inv=pow(n, -1, fr.modulus),
        generator=fr.generator,
        generator_inv=pow(fr.generator, -1, fr.modulus),
    )


def ntt_def(fr: FieldParams, x: Sequence[int], inverse: bool = False, coset: bool = False) -> List[int]:
    """O(n^2) definition.  fft: X[k]=sum x_j w^{jk}; coset_fft: x_j*=g^j first;
    ifft: x[j] = n^-1 sum X_k w^{-jk}; coset_ifft: ifft then *= g^{-j}."""
    n = len(x)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    p = fr.modulus
    d = domain_constants(fr, log_n)
    x = [v % p for v in x]
    if not inverse:
        if coset:
            x = [v * pow(d["generator"], j, p) % p for j, v in enumerate(x)]
        w = d["group_gen"]
        return [sum(x[j] * pow(w, j * k, p) for j in range(n)) % p for k in range(n)]
    w = d["group_gen_inv"]
    y = [sum(x[j] * pow(w, j * k, p) for j in range(n)) * d["size_inv"] % p for k in range(n)]
    if coset:
        y = [v * pow(d["generator_inv"], j, p) % p for j, v in enumerate(y)]
    return y


def _rec(x: List[int], w: int, p: int) -> List[int]:
    n = len(x)
    if n == 1:
        return x
    e = _rec(x[0::2], w * w % p, p)
    o = _rec(x[1::2], w * w % p, p)
    out = [0] * n
    t = 1
    h = n // 2
    for k in range(h):
        v = t * o[k] % p
        out[k] = (e[k] + v) % p
        out[k + h] = (e[k] - v) % p
        t = t * w % p
    return out


def ntt_fast(fr: FieldParams, x: Sequence[int], inverse: bool = False, coset: bool = False) -> List[int]:
    n = len(x)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    p = fr.modulus
    d = domain_constants(fr, log_n)
    x = [v % p for v in x]
    if not inverse:
        if coset:
            g, t = d["generator"], 1
            for j in range(n):
                x[j] = x[j] * t % p
                t = t * g % p
        return _rec(x, d["group_gen"], p)
    y = _rec(x, d["group_gen_inv"], p)
    y = [v * d["size_inv"] % p for v in y]
    if coset:
        g, t = d["generator_inv"], 1
        for j in range(n):
            y[j] = y[j] * t % p
            t = t * g % p
    return y


def horner_eval(fr: FieldParams, coeffs: Sequence[int], point: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * point + c) % fr.modulus
    return acc


# ------------------------------------------------------------------ deterministic inputs
def rand_scalars(fr: FieldParams, n: int, seed: int, kind: str = "uniform") -> List[int]:
    """kind 'uniform' | 'witness' (45 % zero, 45 % one, 10 % uniform) | 'small'."""
    rng = random.Random(seed)
    out = []
    for _ in range(n):
        if kind == "uniform":
            out.append(rng.randrange(fr.modulus))
        elif kind == "witness":
            u = rng.random()
            out.append(0 if u < 0.45 else (1 if u < 0.9 else rng.randrange(fr.modulus)))
        elif kind == "small":
            out.append(rng.randrange(1 << 16))
        else:
            raise ValueError(kind)
    return out

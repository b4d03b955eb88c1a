"""TEST INFRASTRUCTURE ONLY (oracle) -- never imported by the product path.

ctypes binding + build recipe for oracle/cpp/zkm_oracle.cpp, the C++ restatement of the
arkworks 0.3.0 MSM / radix-2 FFT algorithms (see that file's header for the upstream paths
and for the PARITY UNPINNED statement).  Arrays are numpy uint64 in arkworks' own formats:

  field element   : (..., L64) little-endian u64 limbs, Montgomery (value * 2^(64 L64) mod p)
  scalar          : (n, S64) canonical little-endian u64 limbs (Fr::into_repr()); S64 = 4, or 6 for BW6-761
  G1 affine point : (n, 2, L64)       x, y        + separate uint8 infinity flags
  G2 affine point : (n, 2, 2, L64)    x.c0, x.c1, y.c0, y.c1   (BW6-761: (n, 2, L64) like G1 -- its G2 is over Fq)
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from .py.params import BLS12_381, BN254, BW6_761, CurveParams

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "cpp", "zkm_oracle.cpp")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libzkm_oracle.so")

CURVES = {0: BLS12_381, 1: BN254, 2: BW6_761}


def build(force: bool = False) -> str:
    """g++ -O3 -fopenmp; x86-64-v3 + ADX so the .so built here runs on the GPU box's host CPU."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-O3", "-march=x86-64-v3", "-madx", "-fopenmp", "-std=c++17", "-shared", "-fPIC",
           "-o", LIB, SRC]
    subprocess.check_call(cmd)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build()
        L = ctypes.CDLL(path)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        L.orc_msm.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u8p, u64p, ctypes.c_size_t, u64p, u8p, ctypes.c_int]
        L.orc_ntt.argtypes = [ctypes.c_int, u64p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.orc_domain.argtypes = [ctypes.c_int, ctypes.c_int, u64p]
        L.orc_progression.argtypes = [ctypes.c_int, ctypes.c_int, u64p, ctypes.c_uint64, ctypes.c_uint64,
                                      ctypes.c_size_t, u64p]
        L.orc_field_op.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u64p, u64p]
        L.orc_msm_window_bits.argtypes = [ctypes.c_size_t]
        L.orc_witness_map.argtypes = [ctypes.c_int, u64p, u64p, u64p, ctypes.c_int, u64p, ctypes.c_int]
        L.orc_fr_into_repr.argtypes = [ctypes.c_int, u64p, u64p, ctypes.c_size_t]
        _lib = L
    return _lib


def _p64(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))


def _p8(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))


def coord_words(curve_id: int, group: int) -> int:
    c = CURVES[curve_id]
    return c.fq.limbs64 * c.coord_degree(group)


def fr_words(curve_id: int) -> int:
    """u64 words of an Fr element / canonical scalar: 4, or 6 for BW6-761 (BigInteger384)."""
    return CURVES[curve_id].fr.limbs64


def msm(curve_id: int, group: int, bases: np.ndarray, scalars: np.ndarray, infinity=None, threads: int = 0):
    """arkworks-0.3.0 multi_scalar_mul, result normalised with into_affine().
    Returns (xy uint64 array of 2*W words, is_infinity)."""
    W = coord_words(curve_id, group)
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 2 * W)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, fr_words(curve_id))
    n = min(len(bases), len(scalars))
    out = np.zeros(2 * W, dtype=np.uint64)
    oinf = np.zeros(1, dtype=np.uint8)
    infp = None
    if infinity is not None:
        infinity = np.ascontiguousarray(infinity, dtype=np.uint8)
        infp = _p8(infinity)
    rc = lib().orc_msm(curve_id, group, _p64(bases), infp, _p64(scalars), n, _p64(out), _p8(oinf), threads)
    if rc:
        raise RuntimeError("orc_msm rc=%d" % rc)
    return out, bool(oinf[0])


def ntt(curve_id: int, data: np.ndarray, inverse: bool = False, coset: bool = False, threads: int = 0) -> np.ndarray:
    """Radix2EvaluationDomain::{fft,ifft,coset_fft,coset_ifft}_in_place on a copy of `data` (n, 4)."""
    x = np.array(data, dtype=np.uint64, order="C").reshape(-1, fr_words(curve_id))
    n = len(x)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    rc = lib().orc_ntt(curve_id, _p64(x), log_n, int(inverse), int(coset), threads)
    if rc:
        raise RuntimeError("orc_ntt rc=%d" % rc)
    return x


def witness_map(curve_id: int, a: np.ndarray, b: np.ndarray, c: np.ndarray, threads: int = 0) -> np.ndarray:
    """ark-groth16 0.3.0 R1CStoQAP::witness_map after the matrix-vector products: returns h (n, 4)."""
    S = fr_words(curve_id)
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, S)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, S)
    c = np.ascontiguousarray(c, dtype=np.uint64).reshape(-1, S)
    n = len(a)
    log_n = n.bit_length() - 1
    assert 1 << log_n == n and len(b) == n and len(c) == n
    h = np.zeros_like(a)
    rc = lib().orc_witness_map(curve_id, _p64(a), _p64(b), _p64(c), log_n, _p64(h), threads)
    if rc:
        raise RuntimeError("orc_witness_map rc=%d" % rc)
    return h


def domain(curve_id: int, log_n: int) -> dict:
    out = np.zeros((5, fr_words(curve_id)), dtype=np.uint64)
    rc = lib().orc_domain(curve_id, log_n, _p64(out))
    if rc:
        raise ValueError("log_n exceeds two-adicity")
    names = ["group_gen", "group_gen_inv", "size_inv", "generator", "generator_inv"]
    return {k: out[i].copy() for i, k in enumerate(names)}


def generator_mont(curve_id: int, group: int) -> np.ndarray:
    c: CurveParams = CURVES[curve_id]
    L = c.fq.limbs64
    def fe(v):
        m = c.fq.to_mont(v)
        return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(L)]
    if group == 1:
        return np.array(fe(c.g1[0]) + fe(c.g1[1]), dtype=np.uint64)
    if c.g2_over_fq:
        return np.array(fe(c.g2[0]) + fe(c.g2[1]), dtype=np.uint64)
    (x0, x1), (y0, y1) = c.g2
    return np.array(fe(x0) + fe(x1) + fe(y0) + fe(y1), dtype=np.uint64)


def progression(curve_id: int, group: int, a0: int, d: int, n: int) -> np.ndarray:
    """Bases with known discrete logs: P_i = (a0 + i d) G, affine Montgomery, shape (n, 2 W)."""
    W = coord_words(curve_id, group)
    gen = generator_mont(curve_id, group)
    out = np.zeros((n, 2 * W), dtype=np.uint64)
    rc = lib().orc_progression(curve_id, group, _p64(gen), a0, d, n, _p64(out))
    if rc:
        raise RuntimeError("orc_progression rc=%d" % rc)
    return out


def field_op(field: int, op: int, a: np.ndarray, b: np.ndarray = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = a if b is None else np.ascontiguousarray(b, dtype=np.uint64)
    out = np.zeros_like(a)
    rc = lib().orc_field_op(field, op, _p64(a), _p64(b), _p64(out))
    if rc:
        raise RuntimeError("orc_field_op rc=%d" % rc)
    return out


def fr_into_repr(curve_id: int, a: np.ndarray) -> np.ndarray:
    """Fr::into_repr() of every element (Montgomery -> canonical), shape preserved."""
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, fr_words(curve_id))
    out = np.zeros_like(a)
    rc = lib().orc_fr_into_repr(curve_id, _p64(a), _p64(out), len(a))
    if rc:
        raise RuntimeError("orc_fr_into_repr rc=%d" % rc)
    return out


def msm_window_bits(n: int) -> int:
    return lib().orc_msm_window_bits(n)


# ------------------------------------------------------------------ numpy <-> python-int helpers
def ints_to_limbs(vals, limbs64: int) -> np.ndarray:
    out = np.zeros((len(vals), limbs64), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(limbs64):
            out[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    return out


def limbs_to_ints(arr: np.ndarray):
    arr = np.asarray(arr, dtype=np.uint64)
    flat = arr.reshape(-1, arr.shape[-1])
    return [sum(int(flat[i, j]) << (64 * j) for j in range(flat.shape[1])) for i in range(flat.shape[0])]


def random_scalars(curve_id: int, n: int, seed: int, kind: str = "uniform") -> np.ndarray:
    """Canonical scalars (n, 4) u64.  'uniform': 4 x u64 from PCG64, top bits masked to the modulus
    width, rejected while >= r (the shape of ark-ff's Fr::rand); 'witness': 45 % zero, 45 % one,
    10 % uniform; 'small': < 2^16."""
    fr = CURVES[curve_id].fr
    S = fr.limbs64
    rng = np.random.Generator(np.random.PCG64(seed))
    r_limbs = np.array([(fr.modulus >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(S)], dtype=np.uint64)
    top_mask = np.uint64((1 << (fr.bits - 64 * (S - 1))) - 1)

    def uniform(m):
        out = rng.integers(0, 1 << 64, size=(m, S), dtype=np.uint64)
        out[:, S - 1] &= top_mask
        while True:
            ge = np.zeros(m, dtype=bool)
            decided = np.zeros(m, dtype=bool)
            for j in range(S - 1, -1, -1):
                gt = (out[:, j] > r_limbs[j]) & ~decided
                lt = (out[:, j] < r_limbs[j]) & ~decided
                ge |= gt
                decided |= gt | lt
            ge |= ~decided  # equal to r
            k = int(ge.sum())
            if k == 0:
                return out
            fresh = rng.integers(0, 1 << 64, size=(k, S), dtype=np.uint64)
            fresh[:, S - 1] &= top_mask
            out[ge] = fresh

    if kind == "uniform":
        return uniform(n)
    if kind == "small":
        out = np.zeros((n, S), dtype=np.uint64)
        out[:, 0] = rng.integers(0, 1 << 16, size=n, dtype=np.uint64)
        return out
    if kind == "witness":
        out = uniform(n)
        u = rng.random(n)
        out[u < 0.45] = 0
        ones = (u >= 0.45) & (u < 0.9)
        out[ones] = np.array([1] + [0] * (S - 1), dtype=np.uint64)
        return out
    raise ValueError(kind)


def random_field_elements(curve_id: int, n: int, seed: int) -> np.ndarray:
    """Uniform Fr elements; like ark-ff's rand the sampled limbs ARE the Montgomery representation."""
    return random_scalars(curve_id, n, seed, "uniform")

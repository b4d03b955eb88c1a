"""Host-side mirror of the hot part of ark-groth16 0.3.0's prover.

`witness_map` replaces the FFT section of `R1CStoQAP::witness_map` (src/r1cs_to_qap.rs; reached from
/root/reference/benches/groth16.rs:115): the caller (unchanged Rust/host code) evaluates the R1CS
matrices on the assignment -- a[i] = <A_i, z>, b[i] = <B_i, z>, c[i] = <C_i, z>, a[nc + j] = z_j --
and this runs the seven NTTs and the pointwise step on the GPU without leaving HBM.

`ProvingKeyMSMs` holds the five query vectors of a Groth16 proving key registered once on the device
(`pk` is reused for every proof, benches/groth16.rs:107-115) and runs the five MSMs of `create_proof`
(src/prover.rs: h_query, l_query, a_query, b_g1_query on G1, b_g2_query on G2).

`ProvingKey` + `create_proof` mirror `create_proof_with_reduction_and_matrices` from the witness map to the
`Proof { a, b, c }` (final assembly included: `calculate_coeff`, g_c = s g_a + r g1_b - rs delta + l_aux + h_acc),
and `serialize.Proof.serialize()` gives the 192-byte (BLS12-381) arkworks wire format.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .msm import RegisteredBases, AffinePoint, VariableBaseMSM, FR_WORDS, coord_words, _curve_id
from .serialize import Proof, FR_MODULUS


def witness_map(a, b, c, curve="bls12_381") -> np.ndarray:
    """h = coset_ifft((coset_fft(ifft a) * coset_fft(ifft b) - coset_fft(ifft c)) / (g^n - 1)).
    a, b, c: (n, 4) uint64 Montgomery evaluations over the domain, n a power of two."""
    cid = _curve_id(curve)
    S = _lib.FR_WORDS[cid]
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, S)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, S)
    c = np.ascontiguousarray(c, dtype=np.uint64).reshape(-1, S)
    n = len(a)
    log_n = n.bit_length() - 1
    if n == 0 or (1 << log_n) != n or len(b) != n or len(c) != n:
        raise ValueError("a, b, c must have the same power-of-two length")
    h = np.zeros_like(a)
    p = lambda x: ctypes.c_void_p(x.ctypes.data)
    _lib.check(_lib.lib().zkm_witness_map(cid, p(a), p(b), p(c), log_n, p(h)))
    return h


class ProvingKeyMSMs:
    """The five MSM bases of a Groth16 proving key, resident in HBM."""

    def __init__(self, curve, h_query, l_query, a_query, b_g1_query, b_g2_query, infinity=None, precompute=True,
                 g2_device: int | None = None):
        """g2_device: index of the initialised GPU that holds the b_g2 query (the G2 MSM runs on its own GPU
        next to the G1 ones); default: device 1 when the process initialised more than one GPU, else the primary."""
        cid = _curve_id(curve)
        infinity = infinity or {}
        if g2_device is None and _lib.initialised_devices() > 1:
            g2_device = 1
        kw = {"precompute": bool(precompute)}
        self.h = RegisteredBases(cid, 1, h_query, infinity.get("h"), **kw)
        self.l = RegisteredBases(cid, 1, l_query, infinity.get("l"), **kw)
        self.a = RegisteredBases(cid, 1, a_query, infinity.get("a"), **kw)
        self.b_g1 = RegisteredBases(cid, 1, b_g1_query, infinity.get("b_g1"), **kw)
        self.b_g2 = RegisteredBases(cid, 2, b_g2_query, infinity.get("b_g2"), device=g2_device, **kw)

    def prove_msms(self, h_scalars, aux_scalars, full_scalars) -> dict:
        """h_acc, l_acc and the MSM parts of g_a, g1_b, g2_b (calculate_coeff's `acc`).  Upstream runs the
        five MSMs one after another; here they are issued from five host threads, each borrowing its own
        library lane, so the G2 MSM and the serial tails of the G1 ones overlap on the GPU."""
        from concurrent.futures import ThreadPoolExecutor
        jobs = {
            "h_acc": (self.h, h_scalars), "l_acc": (self.l, aux_scalars), "a_acc": (self.a, full_scalars),
            "b_g1_acc": (self.b_g1, full_scalars), "b_g2_acc": (self.b_g2, full_scalars),
        }
        with ThreadPoolExecutor(max_workers=5) as ex:
            futs = {k: ex.submit(reg.msm, sc) for k, (reg, sc) in jobs.items()}
            return {k: f.result() for k, f in futs.items()}

    def release(self):
        for r in (self.h, self.l, self.a, self.b_g1, self.b_g2):
            r.release()


def _scalar_limbs(curve: int, v: int) -> np.ndarray:
    S = FR_WORDS[curve]
    return np.array([(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(S)], dtype=np.uint64)


class ProvingKey:
    """ark_groth16::ProvingKey: the verifying-key elements the prover touches plus the five query vectors.
    `a_query`, `b_g1_query`, `b_g2_query` are the FULL vectors (entry 0 belongs to the constant-one variable and is
    added outside the MSM, as `calculate_coeff` does); the remaining entries are registered on the device once."""

    def __init__(self, curve, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, a_query, b_g1_query, b_g2_query,
                 h_query, l_query, infinity=None, precompute=True):
        self.curve = _curve_id(curve)
        W1, W2 = coord_words(self.curve, 1), coord_words(self.curve, 2)
        as1 = lambda a: np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 2 * W1)
        as2 = lambda a: np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 2 * W2)
        self.alpha_g1, self.beta_g1, self.delta_g1 = as1(alpha_g1)[0], as1(beta_g1)[0], as1(delta_g1)[0]
        self.beta_g2, self.delta_g2 = as2(beta_g2)[0], as2(delta_g2)[0]
        a_query, b_g1_query, b_g2_query = as1(a_query), as1(b_g1_query), as2(b_g2_query)
        infinity = dict(infinity or {})
        flags = {k: np.ascontiguousarray(infinity[k], dtype=np.uint8) if infinity.get(k) is not None else None
                 for k in ("a", "b_g1", "b_g2", "h", "l")}
        self.first = {"a": (a_query[0], bool(flags["a"][0]) if flags["a"] is not None else False),
                      "b_g1": (b_g1_query[0], bool(flags["b_g1"][0]) if flags["b_g1"] is not None else False),
                      "b_g2": (b_g2_query[0], bool(flags["b_g2"][0]) if flags["b_g2"] is not None else False)}
        tail = {k: (flags[k][1:] if flags[k] is not None else None) for k in ("a", "b_g1", "b_g2")}
        self.msms = ProvingKeyMSMs(self.curve, h_query, l_query, a_query[1:], b_g1_query[1:], b_g2_query[1:],
                                   infinity={"h": flags["h"], "l": flags["l"], "a": tail["a"], "b_g1": tail["b_g1"],
                                             "b_g2": tail["b_g2"]}, precompute=precompute)

    def release(self):
        self.msms.release()


def _lincomb(curve: int, group: int, terms) -> AffinePoint:
    """sum of k_i * P_i for a handful of (xy limbs, infinity, int scalar) terms -- a tiny MSM through the same C ABI."""
    xy = np.stack([t[0] for t in terms])
    inf = np.array([1 if t[1] else 0 for t in terms], dtype=np.uint8)
    sc = np.stack([_scalar_limbs(curve, t[2] % FR_MODULUS[curve]) for t in terms])
    return VariableBaseMSM.multi_scalar_mul(xy, sc, curve=curve, group=group, infinity=inf)


def create_proof(pk: ProvingKey, r: int, s: int, a, b, c, input_assignment, aux_assignment) -> Proof:
    """ark_groth16::create_proof_with_reduction_and_matrices (ark-groth16 0.3.0 src/prover.rs) after constraint
    synthesis: `a`, `b`, `c` are the evaluation vectors handed to the witness map (Montgomery Fr, domain size n),
    `input_assignment` (without the leading one) and `aux_assignment` the canonical (`into_repr`) assignments,
    r and s the prover's blinding scalars."""
    from .kzg import KZG10
    curve = pk.curve
    S = FR_WORDS[curve]
    h = witness_map(a, b, c, curve=curve)
    n = len(h)
    inputs = np.ascontiguousarray(input_assignment, dtype=np.uint64).reshape(-1, S)
    aux = np.ascontiguousarray(aux_assignment, dtype=np.uint64).reshape(-1, S)
    full = np.concatenate([inputs, aux])
    # h_acc = MSM(h_query, h[..n-1].into_repr()): into_repr + MSM on the device (the commit entry point does both)
    h_acc = KZG10.commit(pk.msms.h, h[:n - 1])
    l_acc = pk.msms.l.msm(aux)
    a_acc = pk.msms.a.msm(full)
    b1_acc = pk.msms.b_g1.msm(full)
    b2_acc = pk.msms.b_g2.msm(full)
    one = 1
    # calculate_coeff(initial, query, vk_param, assignment) = initial + query[0] + MSM(query[1..], assignment) + vk_param
    g_a = _lincomb(curve, 1, [(pk.delta_g1, False, r), (pk.first["a"][0], pk.first["a"][1], one),
                              (a_acc.xy, a_acc.infinity, one), (pk.alpha_g1, False, one)])
    g1_b = _lincomb(curve, 1, [(pk.delta_g1, False, s), (pk.first["b_g1"][0], pk.first["b_g1"][1], one),
                               (b1_acc.xy, b1_acc.infinity, one), (pk.beta_g1, False, one)])
    g2_b = _lincomb(curve, 2, [(pk.delta_g2, False, s), (pk.first["b_g2"][0], pk.first["b_g2"][1], one),
                               (b2_acc.xy, b2_acc.infinity, one), (pk.beta_g2, False, one)])
    # g_c = s g_a + r g1_b - (r s) delta_g1 + l_aux_acc + h_acc
    g_c = _lincomb(curve, 1, [(g_a.xy, g_a.infinity, s), (g1_b.xy, g1_b.infinity, r), (pk.delta_g1, False, -(r * s)),
                              (l_acc.xy, l_acc.infinity, one), (h_acc.xy, h_acc.infinity, one)])
    return Proof(a=g_a, b=g2_b, c=g_c)

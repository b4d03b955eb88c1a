"""GPU parity at BASELINE.json's full sizes (configs[1] MSM sweep 2^16-2^26, configs[2] NTT sweep 2^16-2^26).

* NTT 2^20 and 2^24: every output byte against the CPU oracle (oracle/cpp: ark-poly's io_helper / oi_helper /
  derange restated), all four modes, through the host C-ABI call `zkm_ntt`.
* NTT 2^25 / 2^26 (the natural THREE-pass plan, k > 24): device-resident; ifft(fft(x)) == x, coset round trip,
  and outputs of a short polynomial checked by Horner evaluation with exact big-int arithmetic.
* MSM 2^24 / 2^25 / 2^26 (BLS12-381, BN254 G1) and 2^20 / 2^22 (G2, BW6-761): the known-discrete-log identity
  sum_i s_i (a0 + i d) G == (sum_i s_i (a0 + i d) mod r) G; the right-hand side is exact big-int arithmetic
  (oracle/py), the bases come from the device generator whose head is pinned against the oracle's progression.
Integer work: the bar is byte equality."""
import ctypes

import numpy as np
import pytest

from oracle import capi, checks
from oracle.py import exact
from oracle.py.params import BLS12_381, BN254, BW6_761

pytestmark = pytest.mark.gpu

MODES = [(False, False), (True, False), (False, True), (True, True)]


@pytest.fixture(scope="module")
def zkm():
    import zkmember_b200 as z
    z.init(0)
    return z


# ----------------------------------------------------------------------------------------- NTT, full byte compare
@pytest.mark.parametrize("curve,log_n", [(BLS12_381, 20), (BN254, 20), (BW6_761, 20), (BLS12_381, 24), (BN254, 24)],
                         ids=lambda v: getattr(v, "name", str(v)))
def test_ntt_bytes_equal_oracle_large(zkm, curve, log_n):
    n = 1 << log_n
    data = capi.random_field_elements(curve.curve_id, n, seed=0x5EED1000 + log_n)
    dom = zkm.Radix2EvaluationDomain(curve.name, log_n)
    for inverse, coset in MODES:
        want = capi.ntt(curve.curve_id, data, inverse, coset)
        got = data.copy()
        dom._run_in_place(got, inverse, coset)
        assert np.array_equal(got, want), "log_n=%d inverse=%s coset=%s" % (log_n, inverse, coset)
        del want, got


def _device_random_fr(torch, curve, n, seed):
    """Uniform-looking canonical Fr elements generated on the device: the top limb is masked below the modulus'
    top limb, so every element is < r (the sampled limbs ARE the Montgomery representation, as in ark-ff's rand)."""
    S = curve.fr.limbs64
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    x = torch.randint(-(1 << 63), (1 << 63) - 1, (n, S), dtype=torch.int64, device="cuda", generator=g)
    top_bits = (curve.fr.modulus >> (64 * (S - 1))).bit_length() - 1      # strictly below the modulus' top limb
    x[:, S - 1] &= (1 << top_bits) - 1
    torch.cuda.synchronize()      # the library's own streams are non-blocking: inputs must be complete before the call
    return x


@pytest.mark.parametrize("curve,log_n", [(BLS12_381, 25), (BLS12_381, 26), (BN254, 26), (BW6_761, 25)],
                         ids=lambda v: getattr(v, "name", str(v)))
def test_ntt_three_pass_roundtrip_horner(zkm, curve, log_n):
    import torch
    fr = curve.fr
    n = 1 << log_n
    dom = zkm.Radix2EvaluationDomain(curve.name, log_n)
    x = _device_random_fr(torch, curve, n, 0x5EED1000 + log_n)
    ev = torch.empty_like(x)
    back = torch.empty_like(x)
    for coset in (False, True):
        dom.transform_device(x.data_ptr(), ev.data_ptr(), inverse=False, coset=coset)
        dom.transform_device(ev.data_ptr(), back.data_ptr(), inverse=True, coset=coset)
        torch.cuda.synchronize()
        assert torch.equal(back, x), "round trip coset=%s" % coset
    # a short polynomial (4096 coefficients) evaluated over the whole domain: Horner at a few points, both variants
    m = 4096
    xs = torch.zeros_like(x)
    xs[:m] = x[:m]
    torch.cuda.synchronize()
    coeffs = [fr.from_mont(v) for v in capi.limbs_to_ints(x[:m].cpu().numpy().view(np.uint64))]
    d = exact.domain_constants(fr, log_n)
    for coset in (False, True):
        dom.transform_device(xs.data_ptr(), ev.data_ptr(), inverse=False, coset=coset)
        torch.cuda.synchronize()
        for k in (0, 1, 12345, n // 2 + 7, n - 1):
            pt = pow(d["group_gen"], k, fr.modulus)
            if coset:
                pt = pt * fr.generator % fr.modulus
            got = fr.from_mont(capi.limbs_to_ints(ev[k:k + 1].cpu().numpy().view(np.uint64))[0])
            assert got == exact.horner_eval(fr, coeffs, pt), (coset, k)
    del x, ev, back, xs
    torch.cuda.empty_cache()


# ----------------------------------------------------------------------------------------- MSM, known discrete logs
@pytest.mark.parametrize("curve,g,log_n,kind", [
    (BLS12_381, 1, 24, "uniform"), (BLS12_381, 1, 25, "uniform"), (BLS12_381, 1, 26, "uniform"), (BLS12_381, 1, 24, "witness"),
    (BN254, 1, 24, "uniform"), (BN254, 1, 25, "uniform"), (BN254, 1, 26, "uniform"),
    (BLS12_381, 2, 20, "uniform"), (BLS12_381, 2, 22, "uniform"), (BN254, 2, 20, "uniform"), (BN254, 2, 22, "witness"),
    (BW6_761, 1, 20, "uniform"), (BW6_761, 1, 22, "uniform"), (BW6_761, 2, 20, "witness"), (BW6_761, 2, 22, "uniform"),
], ids=lambda v: getattr(v, "name", str(v)))
def test_msm_known_discrete_log_full_size(zkm, curve, g, log_n, kind):
    import torch
    n = 1 << log_n
    a0, d = 0x1234567, 0x89ABCDE
    W = curve.fq.limbs64 * curve.coord_degree(g)
    S = curve.fr.limbs64
    d_bases = torch.empty((n, 2 * W), dtype=torch.int64, device="cuda")
    L = zkm._lib.lib()
    zkm._lib.check(L.zkm_testgen_progression_device(curve.curve_id, g, a0, d, n, ctypes.c_void_p(d_bases.data_ptr()),
                                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    # the generator is pinned against the oracle's chained-addition progression (head and tail)
    assert np.array_equal(d_bases[:48].cpu().numpy().view(np.uint64), capi.progression(curve.curve_id, g, a0, d, 48))
    assert np.array_equal(d_bases[n - 8:].cpu().numpy().view(np.uint64),
                          capi.progression(curve.curve_id, g, a0 + (n - 8) * d, d, 8))
    reg = zkm.RegisteredBases.from_device(curve.name, g, d_bases.data_ptr(), n)
    del d_bases
    d_s = _device_random_fr(torch, curve, n, 0x5EED0000 + log_n + 97 * g)
    if kind == "witness":      # 45 % zero, 45 % one, 10 % uniform
        gen = torch.Generator(device="cuda")
        gen.manual_seed(log_n)
        u = torch.rand(n, device="cuda", generator=gen)
        one = torch.zeros(S, dtype=torch.int64, device="cuda")
        one[0] = 1
        d_s[u < 0.45] = 0
        d_s[(u >= 0.45) & (u < 0.9)] = one
        torch.cuda.synchronize()
    d_out = torch.zeros(2 * W + 1, dtype=torch.int64, device="cuda")
    reg.msm_device(d_s.data_ptr(), n, d_out.data_ptr())
    torch.cuda.synchronize()
    out = d_out.cpu().numpy().view(np.uint64)
    scal = d_s.cpu().numpy().view(np.uint64)
    del d_s
    reg.release()
    torch.cuda.empty_cache()
    k = checks.dlog_sum(scal, a0, d, curve.fr.modulus)
    assert checks.msm_identity_ok(curve.curve_id, g, out, k)

"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI of
libzkm_b200.so (via the ctypes host layer) and is compared bit-for-bit with the CPU oracle -- the
C++ restatement of the arkworks 0.3.0 algorithms (oracle/cpp) and the exact big-int definitions
(oracle/py).  Integer work: the bar is byte equality."""
import ctypes

import numpy as np
import pytest

from oracle import capi
from oracle.py import exact
from oracle.py.params import BLS12_381, BN254, BW6_761

pytestmark = pytest.mark.gpu

CURVES = [BLS12_381, BN254, BW6_761]   # BW6-761: 761-bit Fq (G1 and G2), 377-bit Fr, 6-word scalars
MODES = [(False, False), (True, False), (False, True), (True, True)]


@pytest.fixture(scope="module")
def zkm():
    import zkmember_b200 as z
    z.init(0)
    return z


# ----------------------------------------------------------------------------------------- NTT
@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 18])
def test_ntt_matches_oracle(zkm, curve, log_n):
    n = 1 << log_n
    data = capi.random_field_elements(curve.curve_id, n, seed=0x5EED1000 + log_n)
    dom = zkm.Radix2EvaluationDomain(curve.name, log_n)
    for inverse, coset in MODES:
        want = capi.ntt(curve.curve_id, data, inverse, coset)
        fn = {(False, False): dom.fft, (True, False): dom.ifft, (False, True): dom.coset_fft, (True, True): dom.coset_ifft}[(inverse, coset)]
        got = fn(data)
        assert np.array_equal(got, want), "log_n=%d inverse=%s coset=%s" % (log_n, inverse, coset)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("radix", [6, 8, 11, 12])
def test_ntt_pass_plans(zkm, curve, radix):
    """Different pass decompositions (1, 2, 3 passes; tiles up to 2^12) give the same bytes."""
    zkm.set_option("ntt_max_radix_log", radix)
    try:
        for log_n in (12, 14, 17):
            data = capi.random_field_elements(curve.curve_id, 1 << log_n, seed=77 + log_n)
            dom = zkm.Radix2EvaluationDomain(curve.name, log_n)
            assert np.array_equal(dom.fft(data), capi.ntt(curve.curve_id, data))
            assert np.array_equal(dom.coset_ifft(data), capi.ntt(curve.curve_id, data, True, True))
    finally:
        zkm.set_option("ntt_max_radix_log", 10)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_ntt_zero_padding_and_definition(zkm, curve):
    """fft() resizes short inputs with zeros like upstream; outputs equal the O(n^2) definition."""
    fr = curve.fr
    x = [3, 1, 4, 1, 5]
    dom = zkm.Radix2EvaluationDomain.new(len(x), curve.name)
    assert dom.size == 8
    data = capi.ints_to_limbs([fr.to_mont(v) for v in x], fr.limbs64)
    got = capi.limbs_to_ints(dom.fft(data))
    want = exact.ntt_def(fr, x + [0, 0, 0])
    assert got == [fr.to_mont(v) for v in want]
    got = capi.limbs_to_ints(dom.coset_fft(data))
    assert got == [fr.to_mont(v) for v in exact.ntt_def(fr, x + [0, 0, 0], coset=True)]


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_ntt_large_roundtrip_linearity_horner(zkm, curve):
    """Size-independent properties at 2^22: ifft(fft(x)) == x, coset round trip, linearity, and Horner
    spot checks of a few outputs with exact big-int arithmetic."""
    fr = curve.fr
    log_n = 22
    n = 1 << log_n
    dom = zkm.Radix2EvaluationDomain(curve.name, log_n)
    x = capi.random_field_elements(curve.curve_id, n, seed=0x5EED1000 + log_n)
    ev = dom.fft(x)
    assert np.array_equal(dom.ifft(ev), x)
    cev = dom.coset_fft(x)
    assert np.array_equal(dom.coset_ifft(cev), x)
    coeffs = [fr.from_mont(v) for v in capi.limbs_to_ints(x[:4096])]
    # fft of the truncated polynomial, checked by Horner at three domain points
    xs = np.zeros_like(x)
    xs[:4096] = x[:4096]
    evs = dom.fft(xs)
    d = exact.domain_constants(fr, log_n)
    for k in (1, 12345, n - 1):
        pt = pow(d["group_gen"], k, fr.modulus)
        assert fr.from_mont(capi.limbs_to_ints(evs[k:k + 1])[0]) == exact.horner_eval(fr, coeffs, pt)
    # linearity: fft(x) == fft(xs) + fft(x - xs), spot-checked on a slice with exact arithmetic
    xr = x.copy()
    xr[:4096] = 0
    evr = dom.fft(xr)
    for k in (0, 7, n // 3):
        a = fr.from_mont(capi.limbs_to_ints(evs[k:k + 1])[0])
        b = fr.from_mont(capi.limbs_to_ints(evr[k:k + 1])[0])
        assert (a + b) % fr.modulus == fr.from_mont(capi.limbs_to_ints(ev[k:k + 1])[0])


def test_ntt_domain_errors(zkm):
    assert zkm.Radix2EvaluationDomain.new((1 << 28) + 1, "bn254") is None      # upstream: new() -> None
    assert zkm.Radix2EvaluationDomain.new(1 << 28, "bn254") is not None
    with pytest.raises(zkm.DomainError):
        zkm.Radix2EvaluationDomain("bn254", 29)
    for curve in CURVES:
        for log_n in (0, 5, 16, curve.fr.two_adicity):
            d = zkm.Radix2EvaluationDomain(curve.name, log_n)
            e = exact.domain_constants(curve.fr, log_n)
            for name in ("group_gen", "group_gen_inv", "size_inv", "generator_inv"):
                assert capi.limbs_to_ints(getattr(d, name)[None, :])[0] == curve.fr.to_mont(e[name]), (log_n, name)


# ----------------------------------------------------------------------------------------- MSM
def _check_point(curve, g, got, want_xy, want_inf):
    assert got.infinity == bool(want_inf)
    assert np.array_equal(got.xy, want_xy), (got.xy, want_xy)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("g", [1, 2])
@pytest.mark.parametrize("n,kind", [(0, "uniform"), (1, "uniform"), (2, "uniform"), (7, "uniform"), (33, "small"),
                                    (100, "witness"), (1000, "uniform"), (4096, "uniform"), (5000, "witness")])
def test_msm_matches_oracle(zkm, curve, g, n, kind):
    bases = capi.progression(curve.curve_id, g, 11 + n, 7, n)
    scal = capi.random_scalars(curve.curve_id, n, seed=n * 10 + g, kind=kind)
    inf = np.zeros(n, dtype=np.uint8)
    if n >= 7:
        inf[2] = 1                    # point at infinity among the bases
        bases[4] = bases[3]           # repeated point
        W = bases.shape[1] // 2
        G = exact.Group(curve, g)
        P5 = exact.point_from_bytes(curve, g, bases[5].tobytes(), 0)
        b, _ = exact.point_to_bytes(curve, g, G.neg(P5))
        bases[6] = np.frombuffer(b, dtype=np.uint64)   # P and -P with equal scalars cancel
        scal[6] = scal[5]
    want_xy, want_inf = capi.msm(curve.curve_id, g, bases, scal, inf)
    got = zkm.VariableBaseMSM.multi_scalar_mul(bases, scal, curve=curve.name, group=g, infinity=inf)
    _check_point(curve, g, got, want_xy, want_inf)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("kind", ["uniform", "witness"])
def test_msm_g1_2p16_matches_oracle(zkm, curve, kind):
    n = 1 << 16
    bases = capi.progression(curve.curve_id, 1, 0xABCDEF, 0x1357, n)
    scal = capi.random_scalars(curve.curve_id, n, seed=0x5EED0000 + 16, kind=kind)
    inf = np.zeros(n, dtype=np.uint8)
    rng = np.random.default_rng(5)
    inf[rng.integers(0, n, 16)] = 1
    dup = rng.integers(1, n, 16)
    bases[dup] = bases[dup - 1]
    want_xy, want_inf = capi.msm(curve.curve_id, 1, bases, scal, inf)
    got = zkm.VariableBaseMSM.multi_scalar_mul(bases, scal, curve=curve.name, group=1, infinity=inf)
    _check_point(curve, 1, got, want_xy, want_inf)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("c_bits", [4, 9, 13, 16])
def test_msm_window_sizes_agree(zkm, curve, c_bits):
    """Any window size / chunk length gives the same affine point (the result is unique)."""
    n = 3000
    bases = capi.progression(curve.curve_id, 1, 99, 5, n)
    scal = capi.random_scalars(curve.curve_id, n, seed=c_bits, kind="uniform")
    want_xy, want_inf = capi.msm(curve.curve_id, 1, bases, scal)
    zkm.set_option("msm_window_bits", c_bits)
    zkm.set_option("msm_chunk", 7)
    try:
        got = zkm.VariableBaseMSM.multi_scalar_mul(bases, scal, curve=curve.name, group=1)
    finally:
        zkm.set_option("msm_window_bits", 0)
        zkm.set_option("msm_chunk", 0)
    _check_point(curve, 1, got, want_xy, want_inf)


def test_msm_all_equal_scalars_deep_fold(zkm):
    """Every point lands in the same bucket of each window: exercises the multi-level fold."""
    curve = BLS12_381
    n = 20000
    bases = capi.progression(curve.curve_id, 1, 1, 1, n)
    scal = np.tile(np.array([[0x0123456789ABCDEF, 0x1111, 0, 0]], dtype=np.uint64), (n, 1))
    want_xy, want_inf = capi.msm(curve.curve_id, 1, bases, scal)
    got = zkm.VariableBaseMSM.multi_scalar_mul(bases, scal, curve=curve.name, group=1)
    _check_point(curve, 1, got, want_xy, want_inf)
    # and the all-ones witness (the scalar==1 fast path upstream)
    scal1 = np.zeros((n, 4), dtype=np.uint64)
    scal1[:, 0] = 1
    want_xy, want_inf = capi.msm(curve.curve_id, 1, bases, scal1)
    got = zkm.VariableBaseMSM.multi_scalar_mul(bases, scal1, curve=curve.name, group=1)
    _check_point(curve, 1, got, want_xy, want_inf)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
def test_msm_registered_slices(zkm, curve):
    """KZG10::commit's powers_of_g[z..] slicing: MSM over a sub-range of registered bases."""
    n = 2048
    bases = capi.progression(curve.curve_id, 1, 3, 11, n)
    reg = zkm.RegisteredBases(curve.name, 1, bases)
    try:
        for off, m in ((0, n), (5, 1000), (n - 1, 1), (100, 0)):
            scal = capi.random_scalars(curve.curve_id, m, seed=off + m)
            want_xy, want_inf = capi.msm(curve.curve_id, 1, bases[off:off + m], scal)
            got = reg.msm(scal, offset=off)
            _check_point(curve, 1, got, want_xy, want_inf)
        with pytest.raises(zkm.ZkmError):
            reg.msm(capi.random_scalars(curve.curve_id, 10, 1), offset=n - 5, n=10)
    finally:
        reg.release()
    with pytest.raises(zkm.ZkmError):
        reg2 = zkm.RegisteredBases(curve.name, 1, bases[:4])
        reg2.release()
        reg2.handle = 12345
        reg2.msm(capi.random_scalars(curve.curve_id, 4, 1))


def test_msm_rejects_non_canonical_scalar(zkm):
    bases = capi.progression(0, 1, 1, 1, 8)
    scal = capi.random_scalars(0, 8, 1)
    scal[3, 3] = 0xFFFFFFFFFFFFFFFF
    with pytest.raises(zkm.ZkmError) as ei:
        zkm.VariableBaseMSM.multi_scalar_mul(bases, scal)
    assert ei.value.code == -5
    # the next call on the same lanes is unaffected
    scal[3, 3] = 0
    want_xy, want_inf = capi.msm(0, 1, bases, scal)
    _check_point(BLS12_381, 1, zkm.VariableBaseMSM.multi_scalar_mul(bases, scal), want_xy, want_inf)


def test_msm_non_canonical_scalar_device_record_and_shards(zkm):
    """The range check happens on the device (the run never waits for the host): *_device entry points report it as
    flag word 2 of the result record, host entry points -- also over several parts / the cache -- as ZKM_ERR_SCALAR_RANGE."""
    import torch
    n = 3000
    bases = capi.progression(0, 1, 5, 3, n)
    scal = capi.random_scalars(0, n, 9)
    bad = scal.copy()
    bad[n - 7, 3] = 1 << 63          # bit 255: not below the 255-bit modulus
    reg = zkm.RegisteredBases("bls12_381", 1, bases)
    try:
        d_out = torch.zeros(13, dtype=torch.int64, device="cuda")
        for s_host, flag in ((bad, 2), (scal, 0)):
            d_s = torch.from_numpy(s_host.view(np.int64)).cuda()
            reg.msm_device(d_s.data_ptr(), n, d_out.data_ptr())
            torch.cuda.synchronize()
            assert int(d_out.cpu().numpy().view(np.uint64)[-1]) == flag
        with pytest.raises(zkm.ZkmError) as ei:
            reg.msm(bad)
        assert ei.value.code == -5
        # the literal multi_scalar_mul signature above the cache threshold (n >= 1024)
        with pytest.raises(zkm.ZkmError) as ei:
            zkm.VariableBaseMSM.multi_scalar_mul(bases, bad)
        assert ei.value.code == -5
        want_xy, want_inf = capi.msm(0, 1, bases, scal)
        _check_point(BLS12_381, 1, reg.msm(scal), want_xy, want_inf)
    finally:
        reg.release()


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("g", [1, 2])
@pytest.mark.parametrize("chunk,kind", [(1, "witness"), (2, "uniform"), (3, "witness"), (16, "witness"), (5, "equal")])
def test_msm_fold_kernels(zkm, curve, g, chunk, kind):
    """Buckets cut into several tasks: 2..8 partial sums go through k_fold_quad, more through k_fold_cta (one CTA per
    bucket: strided sums + shared-memory tree).  Tiny task lengths force both paths on every window; `equal` puts every
    point into one bucket per window (thousands of partial sums in one CTA)."""
    n = 6000 if g == 1 else 2500
    bases = capi.progression(curve.curve_id, g, 3, 5, n)
    if kind == "equal":
        scal = np.tile(capi.random_scalars(curve.curve_id, 1, seed=chunk), (n, 1))
    else:
        scal = capi.random_scalars(curve.curve_id, n, seed=100 + chunk, kind=kind)
    inf = np.zeros(n, dtype=np.uint8)
    inf[::97] = 1
    want_xy, want_inf = capi.msm(curve.curve_id, g, bases, scal, inf)
    zkm.set_option("msm_chunk", chunk)
    try:
        for c_bits in (0, 4, 11):
            zkm.set_option("msm_window_bits", c_bits)
            got = zkm.VariableBaseMSM.multi_scalar_mul(bases, scal, curve=curve.name, group=g, infinity=inf)
            _check_point(curve, g, got, want_xy, want_inf)
    finally:
        zkm.set_option("msm_chunk", 0)
        zkm.set_option("msm_window_bits", 0)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("log_n,kind", [(20, "uniform"), (22, "uniform"), (22, "witness")])
def test_msm_known_discrete_log_large(zkm, curve, log_n, kind):
    """Full-size property: with bases P_i = (a0 + i d) G generated on the device,
    sum s_i P_i == (sum s_i (a0 + i d) mod r) G, checked with exact big-int arithmetic."""
    import torch
    n = 1 << log_n      # 2^22 x 16 windows crosses the threshold of the batched-affine pairwise levels
    a0, d = 0x1234567, 0x89ABCDE
    W = curve.fq.limbs64
    dev = torch.device("cuda:0")
    d_bases = torch.empty((n, 2 * W), dtype=torch.int64, device=dev)
    L = zkm._lib.lib()
    zkm._lib.check(L.zkm_testgen_progression_device(curve.curve_id, 1, a0, d, n, ctypes.c_void_p(d_bases.data_ptr()),
                                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    # the generator itself is pinned against the oracle's chained-addition progression
    head = d_bases[:64].cpu().numpy().view(np.uint64)
    assert np.array_equal(head, capi.progression(curve.curve_id, 1, a0, d, 64))
    reg = zkm.RegisteredBases.from_device(curve.name, 1, d_bases.data_ptr(), n)
    scal = capi.random_scalars(curve.curve_id, n, seed=0x5EED0000 + log_n, kind=kind)
    got = reg.msm(scal)
    reg.release()
    r = curve.fr.modulus
    s_int = scal.astype(object)
    s_vals = sum(s_int[:, j] << (64 * j) for j in range(curve.fr.limbs64))
    idx = np.arange(n, dtype=object)
    k = int(np.sum(s_vals * (a0 + idx * d)) % r)
    G = exact.Group(curve, 1)
    want = G.mul(G.gen, k)
    b, f = exact.point_to_bytes(curve, 1, want)
    assert not got.infinity and got.xy.tobytes() == b


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("g", [1, 2])
def test_points_sum_device(zkm, curve, g):
    """The cross-GPU reduction step: sum of a handful of partial results."""
    import torch
    G = exact.Group(curve, g)
    pts = G.progression(5, 9, 6)
    pts[2] = None
    pts[4] = pts[3]
    W = curve.fq.limbs64 * curve.coord_degree(g)
    rec = np.zeros((len(pts), 2 * W + 1), dtype=np.uint64)
    for i, P in enumerate(pts):
        b, f = exact.point_to_bytes(curve, g, P)
        rec[i, :2 * W] = np.frombuffer(b, dtype=np.uint64)
        rec[i, 2 * W] = f
    d_in = torch.from_numpy(rec.view(np.int64)).cuda()
    d_out = torch.zeros(2 * W + 1, dtype=torch.int64, device="cuda")
    L = zkm._lib.lib()
    zkm._lib.check(L.zkm_points_sum_device(curve.curve_id, g, ctypes.c_void_p(d_in.data_ptr()), len(pts),
                                            ctypes.c_void_p(d_out.data_ptr()),
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    out = d_out.cpu().numpy().view(np.uint64)
    want = None
    for P in pts:
        want = G.add(want, P)
    b, f = exact.point_to_bytes(curve, g, want)
    assert out[2 * W] == f and out[:2 * W].tobytes() == b


# ------------------------------------------------------------------------------- 8f rows: witness map, precompute
@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("log_n", [3, 10, 15])
def test_witness_map_matches_oracle(zkm, curve, log_n):
    from zkmember_b200.groth16 import witness_map
    n = 1 << log_n
    a = capi.random_field_elements(curve.curve_id, n, seed=1 + log_n)
    b = capi.random_field_elements(curve.curve_id, n, seed=2 + log_n)
    c = capi.random_field_elements(curve.curve_id, n, seed=3 + log_n)
    want = capi.witness_map(curve.curve_id, a, b, c)
    got = witness_map(a, b, c, curve=curve.name)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("g", [1, 2])
def test_msm_precomputed_window_multiples(zkm, curve, g):
    """Registrations with precomputed window multiples give the same point, incl. offsets, infinity, +-P."""
    n = 3000 if g == 1 else 600
    bases = capi.progression(curve.curve_id, g, 21, 13, n)
    inf = np.zeros(n, dtype=np.uint8)
    inf[[5, 77]] = 1
    bases[9] = bases[8]
    zkm.set_option("msm_precompute", 1)
    try:
        reg = zkm.RegisteredBases(curve.name, g, bases, inf)
    finally:
        zkm.set_option("msm_precompute", 0)
    try:
        for kind, off, m in (("uniform", 0, n), ("witness", 0, n), ("uniform", 100, 333), ("small", n - 1, 1), ("uniform", 7, 0)):
            scal = capi.random_scalars(curve.curve_id, m, seed=off + m + g, kind=kind)
            want_xy, want_inf = capi.msm(curve.curve_id, g, bases[off:off + m], scal, inf[off:off + m])
            got = reg.msm(scal, offset=off)
            _check_point(curve, g, got, want_xy, want_inf)
    finally:
        reg.release()


# ------------------------------------------------------------------------------- batched-affine pairwise levels
@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("g", [1, 2])
@pytest.mark.parametrize("levels", [1, 3, 9])
def test_msm_affine_levels_exceptional_pairs(zkm, curve, g, levels):
    """Forced batched-affine levels on lists built to hit every exceptional pair: P + P (doubling in the
    batch), P + (-P) (identity marker travelling through later levels), odd lists, single entries."""
    G = exact.Group(curve, g)
    base = capi.progression(curve.curve_id, g, 77, 5, 8)
    neg = []
    for i in range(8):
        P = exact.point_from_bytes(curve, g, base[i].tobytes(), 0)
        b, _ = exact.point_to_bytes(curve, g, G.neg(P))
        neg.append(np.frombuffer(b, dtype=np.uint64))
    s = capi.random_scalars(curve.curve_id, 4, seed=levels + g)
    cases = [
        ([base[0], base[0]], [s[0], s[0]]),                                   # P + P in every bucket
        ([base[0], neg[0]], [s[0], s[0]]),                                    # cancels everywhere -> identity
        ([base[0], base[0], base[0]], [s[1], s[1], s[1]]),                    # odd list of equal points
        ([base[0], neg[0], base[1]], [s[1], s[1], s[1]]),                     # identity marker + another point
        ([base[0], base[0], neg[0], neg[0], base[2]], [s[2]] * 5),
        ([base[i % 8] for i in range(37)], [s[3]] * 37),                      # one long list per window, repeats
    ]
    zkm.set_option("msm_affine_levels", levels)
    zkm.set_option("msm_window_bits", 6)
    try:
        for bs, ss in cases:
            bases = np.stack(bs)
            scal = np.stack(ss)
            want_xy, want_inf = capi.msm(curve.curve_id, g, bases, scal)
            got = zkm.VariableBaseMSM.multi_scalar_mul(bases, scal, curve=curve.name, group=g)
            _check_point(curve, g, got, want_xy, want_inf)
    finally:
        zkm.set_option("msm_affine_levels", -1)
        zkm.set_option("msm_window_bits", 0)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("levels,kind", [(1, "uniform"), (2, "witness"), (4, "uniform"), (6, "witness"), (12, "small")])
def test_msm_affine_levels_random(zkm, curve, levels, kind):
    n = 6000
    bases = capi.progression(curve.curve_id, 1, 31, 17, n)
    inf = np.zeros(n, dtype=np.uint8)
    inf[[1, 500]] = 1
    bases[41] = bases[40]
    scal = capi.random_scalars(curve.curve_id, n, seed=levels, kind=kind)
    scal[41] = scal[40]
    want_xy, want_inf = capi.msm(curve.curve_id, 1, bases, scal, inf)
    zkm.set_option("msm_affine_levels", levels)
    zkm.set_option("msm_window_bits", 7)
    try:
        got = zkm.VariableBaseMSM.multi_scalar_mul(bases, scal, curve=curve.name, group=1, infinity=inf)
    finally:
        zkm.set_option("msm_affine_levels", -1)
        zkm.set_option("msm_window_bits", 0)
    _check_point(curve, 1, got, want_xy, want_inf)


# ------------------------------------------------------------------------------- 8f row 3: KZG10 commit
@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("precompute", [False, True])
def test_kzg_commit_matches_oracle(zkm, curve, precompute):
    """KZG10::commit (non-hiding part): leading-zero skip + into_repr + MSM over powers[z..]."""
    from zkmember_b200.kzg import KZG10, Powers
    fr = curve.fr
    n = 1500
    powers = capi.progression(curve.curve_id, 1, 5, 3, n)       # stand-in SRS (any G1 points)
    pw = Powers(curve.name, powers, precompute=precompute)
    fid = {0: 1, 1: 3, 2: 5}[curve.curve_id]
    try:
        for deg, lead in ((n - 1, 0), (999, 7), (10, 11), (0, 0)):
            coeffs = capi.random_field_elements(curve.curve_id, deg + 1, seed=deg + lead)
            coeffs[:min(lead, deg + 1)] = 0
            z = 0
            while z < len(coeffs) and not coeffs[z].any():
                z += 1
            repr_ = np.stack([capi.field_op(fid, 4, coeffs[i]) for i in range(z, len(coeffs))]) if z < len(coeffs) \
                else np.zeros((0, fr.limbs64), dtype=np.uint64)
            want_xy, want_inf = capi.msm(curve.curve_id, 1, powers[z:z + len(repr_)], repr_)
            got = KZG10.commit(pw, coeffs)
            _check_point(curve, 1, got, want_xy, want_inf)
    finally:
        pw.release()


# ------------------------------------------------------------------------------- lanes: concurrent host threads
def test_concurrent_calls_from_threads(zkm):
    """Compute calls borrow independent lanes: MSMs (G1 + G2), NTTs and witness maps issued from several
    host threads at once must all return the bytes of the serial run."""
    from concurrent.futures import ThreadPoolExecutor
    from zkmember_b200.groth16 import witness_map
    curve = BLS12_381
    n = 4096
    b1 = capi.progression(0, 1, 3, 5, n)
    b2 = capi.progression(0, 2, 3, 5, 512)
    regs = [zkm.RegisteredBases("bls12_381", 1, b1), zkm.RegisteredBases("bls12_381", 2, b2)]
    scal = [capi.random_scalars(0, n, seed=s, kind=k) for s, k in ((1, "uniform"), (2, "witness"), (3, "uniform"))]
    x = capi.random_field_elements(0, 1 << 13, seed=9)
    dom = zkm.Radix2EvaluationDomain("bls12_381", 13)
    a, b, c = (capi.random_field_elements(0, 1 << 12, seed=s) for s in (4, 5, 6))
    jobs = [
        lambda: ("m0", regs[0].msm(scal[0]).to_bytes()),
        lambda: ("m1", regs[0].msm(scal[1]).to_bytes()),
        lambda: ("m2", regs[1].msm(scal[2][:512]).to_bytes()),
        lambda: ("f", dom.fft(x).tobytes()),
        lambda: ("ci", dom.coset_ifft(x).tobytes()),
        lambda: ("w", witness_map(a, b, c).tobytes()),
    ]
    want = dict(j() for j in jobs)                      # serial reference
    with ThreadPoolExecutor(max_workers=6) as ex:
        for _ in range(4):                                # several rounds of 12 concurrent calls
            got = [f.result() for f in [ex.submit(j) for j in jobs + jobs]]
            for k, v in got:
                assert v == want[k], k
    xy, inf = capi.msm(0, 1, b1, scal[0])
    assert want["m0"] == xy.tobytes() + bytes([1 if inf else 0])
    for r in regs:
        r.release()


# ------------------------------------------------------------------------------- error behaviour of the C ABI
def test_c_abi_error_codes(zkm):
    """Errors surface as negative codes + message, never as aborts; valid calls keep working afterwards."""
    L = zkm._lib.lib()
    x = np.zeros((8, 4), dtype=np.uint64)
    px = ctypes.c_void_p(x.ctypes.data)
    assert L.zkm_ntt(7, px, 3, 0, 0) == -1 and b"curve" in L.zkm_last_error()          # ZKM_ERR_ARG
    assert L.zkm_ntt(0, ctypes.c_void_p(0), 3, 0, 0) == -1
    assert L.zkm_ntt(1, px, 29, 0, 0) == -4                                              # ZKM_ERR_DOMAIN (BN254: 28)
    assert L.zkm_ntt(0, px, 33, 0, 0) == -4
    out = np.zeros(12, dtype=np.uint64)
    inf = np.zeros(1, dtype=np.uint8)
    po, pi = ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(inf.ctypes.data)
    assert L.zkm_msm_g1(5, px, ctypes.c_void_p(0), px, 1, po, pi) == -1
    assert L.zkm_msm_g1(0, ctypes.c_void_p(0), ctypes.c_void_p(0), px, 1, po, pi) == -1  # null bases with n > 0
    assert L.zkm_msm_g1(0, ctypes.c_void_p(0), ctypes.c_void_p(0), ctypes.c_void_p(0), 0, po, pi) == 0  # empty sum
    assert inf[0] == 1 and not out[:6].any()                                              # identity: x = 0, y = 1 (Montgomery)
    assert capi.limbs_to_ints(out[None, 6:])[0] == BLS12_381.fq.to_mont(1)
    assert L.zkm_msm_registered(987654, 0, px, 1, po, pi) == -6                          # ZKM_ERR_HANDLE
    assert L.zkm_bases_release(987654) == -6
    assert L.zkm_set_option(b"no_such_option", 1) == -1
    assert L.zkm_set_option(b"msm_window_bits", 99) == -1
    assert L.zkm_init(0) == 0                                                             # idempotent
    if L.zkm_device_count() > 1:
        assert L.zkm_init(1) == -1                                                        # one process per GPU
    # the library is still healthy
    dom = zkm.Radix2EvaluationDomain("bls12_381", 3)
    data = capi.random_field_elements(0, 8, seed=1)
    assert np.array_equal(dom.fft(data), capi.ntt(0, data))


def test_msm_truncates_to_shorter_input_like_upstream(zkm):
    """multi_scalar_mul uses min(bases.len(), scalars.len()) pairs."""
    bases = capi.progression(0, 1, 9, 4, 50)
    scal = capi.random_scalars(0, 80, seed=3)
    got = zkm.VariableBaseMSM.multi_scalar_mul(bases, scal)                 # 50 bases, 80 scalars
    want_xy, want_inf = capi.msm(0, 1, bases, scal[:50])
    _check_point(BLS12_381, 1, got, want_xy, want_inf)
    got = zkm.VariableBaseMSM.multi_scalar_mul(bases, scal[:20])            # 50 bases, 20 scalars
    want_xy, want_inf = capi.msm(0, 1, bases[:20], scal[:20])
    _check_point(BLS12_381, 1, got, want_xy, want_inf)


def test_proving_key_msms_concurrent(zkm):
    """The five MSMs of create_proof issued concurrently (zkmember_b200.groth16.ProvingKeyMSMs)."""
    from zkmember_b200.groth16 import ProvingKeyMSMs
    n = 1 << 10
    q = {k: capi.progression(0, 2 if k == "b_g2" else 1, 3 + i, 2 + i, n) for i, k in enumerate(("h", "l", "a", "b_g1", "b_g2"))}
    inf = {"b_g1": (np.arange(n) % 2).astype(np.uint8), "b_g2": (np.arange(n) % 3 == 0).astype(np.uint8)}
    pk = ProvingKeyMSMs("bls12_381", q["h"], q["l"], q["a"], q["b_g1"], q["b_g2"], infinity=inf, precompute=True)
    h = capi.random_scalars(0, n, 1)
    w = capi.random_scalars(0, n, 2, "witness")
    got = pk.prove_msms(h, w, w)
    pk.release()
    for name, key, sc, g in (("h_acc", "h", h, 1), ("l_acc", "l", w, 1), ("a_acc", "a", w, 1), ("b_g1_acc", "b_g1", w, 1),
                             ("b_g2_acc", "b_g2", w, 2)):
        xy, isinf = capi.msm(0, g, q[key], sc, inf.get(key))
        _check_point(BLS12_381, g, got[name], xy, isinf)


def test_msm_batch_registered_device(zkm):
    """zkm_msm_batch_registered_device: five concurrent MSMs in one call == five separate calls."""
    import torch
    from zkmember_b200.msm import msm_batch_device
    n = 2000
    regs, outs, want = [], [], []
    scal = capi.random_scalars(0, n, 5, "witness")
    d_s = torch.from_numpy(scal.view(np.int64)).cuda()
    for i, g in enumerate((1, 1, 2, 1, 1)):
        b = capi.progression(0, g, 7 + i, 3, n)
        regs.append(zkm.RegisteredBases("bls12_381", g, b))
        outs.append(torch.zeros(12 * g + 1, dtype=torch.int64, device="cuda"))
        want.append(capi.msm(0, g, b, scal))
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        msm_batch_device([(regs[i], d_s.data_ptr(), n, outs[i].data_ptr()) for i in range(5)], stream=st.cuda_stream)
        st.synchronize()
    for i in range(5):
        r = outs[i].cpu().numpy().view(np.uint64)
        assert bool(r[-1]) == want[i][1] and np.array_equal(r[:-1], want[i][0]), i
        regs[i].release()

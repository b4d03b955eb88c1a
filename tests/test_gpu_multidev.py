"""One process, several GPUs (SURVEY 8b `zkm_init(device_mask)`, 8e range-sharded MSM with the
partial sums reduced on one device, G2 MSM on its own GPU) -- through the C ABI.

On a box with >= 2 GPUs the module initialises devices [0, 1]; on the one-GPU test box it initialises the SAME GPU
twice (`zkm_init_devices([0, 0])`: two lane sets acting as two devices), which drives exactly the same code --
sharded registrations, per-device threads, peer copies of scalar slices and result records, k_points_sum on the
caller's device.  Also here: the registration cache of the literal multi_scalar_mul(bases, scalars) call and the
profile counters."""
import ctypes

import numpy as np
import pytest

from oracle import capi
from oracle.py.params import BLS12_381, BN254, BW6_761

pytestmark = pytest.mark.gpu

CURVES = [BLS12_381, BN254, BW6_761]


@pytest.fixture(scope="module")
def zkm2():
    import zkmember_b200 as z
    z.shutdown()
    L = z.load()
    devs = [0, 1] if L.zkm_device_count() >= 2 else [0, 0]
    z.init(devs)
    assert z._lib.initialised_devices() == 2
    yield z
    z.shutdown()
    z.init(0)


def _check(got, want_xy, want_inf):
    assert got.infinity == bool(want_inf) and np.array_equal(got.xy, want_xy)


@pytest.mark.parametrize("curve", CURVES, ids=lambda c: c.name)
@pytest.mark.parametrize("g", [1, 2])
@pytest.mark.parametrize("precompute", [False, True])
def test_sharded_registration_host_msm(zkm2, curve, g, precompute):
    n = 3001 if g == 1 else 601                      # odd: the shards differ in length
    bases = capi.progression(curve.curve_id, g, 17, 3, n)
    inf = np.zeros(n, dtype=np.uint8)
    inf[[4, n // 2, n - 1]] = 1
    bases[9] = bases[8]
    reg = zkm2.RegisteredBases(curve.name, g, bases, inf, shard=True, precompute=precompute)
    try:
        for kind, off, m in (("uniform", 0, n), ("witness", 0, n), ("uniform", n // 2 - 7, 20), ("uniform", 3, n // 2 - 3),
                             ("uniform", n // 2, n - n // 2), ("small", n - 1, 1), ("uniform", 11, 0)):
            scal = capi.random_scalars(curve.curve_id, m, seed=off + m + g, kind=kind)
            want_xy, want_inf = capi.msm(curve.curve_id, g, bases[off:off + m], scal, inf[off:off + m])
            _check(reg.msm(scal, offset=off), want_xy, want_inf)
    finally:
        reg.release()


@pytest.mark.parametrize("placement", ["shard", "device1", "device0"])
def test_device_pointer_msm_over_remote_and_sharded_registrations(zkm2, placement):
    import torch
    n = 5000
    bases = capi.progression(0, 1, 5, 7, n)
    kw = {"shard": True} if placement == "shard" else {"device": 1 if placement == "device1" else 0}
    reg = zkm2.RegisteredBases("bls12_381", 1, bases, **kw)
    try:
        scal = capi.random_scalars(0, n, seed=3)
        d_s = torch.from_numpy(scal.view(np.int64)).cuda()
        d_out = torch.zeros(13, dtype=torch.int64, device="cuda")
        torch.cuda.synchronize()
        for off, m in ((0, n), (100, 4000), (n // 2 - 1, 2)):
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                reg.msm_device(d_s.data_ptr() + off * 32, m, d_out.data_ptr(), offset=off, stream=st.cuda_stream)
                st.synchronize()
            rec = d_out.cpu().numpy().view(np.uint64)
            want_xy, want_inf = capi.msm(0, 1, bases[off:off + m], scal[off:off + m])
            assert bool(rec[-1]) == want_inf and np.array_equal(rec[:-1], want_xy), (placement, off, m)
    finally:
        reg.release()


def test_proving_key_with_g2_on_the_second_device(zkm2):
    """create_proof's five MSMs in one batch call: four G1 registrations on device 0, the G2 query on device 1."""
    import torch
    from zkmember_b200.msm import msm_batch_device
    n = 2000
    scal = capi.random_scalars(0, n, 5, "witness")
    d_s = torch.from_numpy(scal.view(np.int64)).cuda()
    regs, outs, want = [], [], []
    for i, (g, dev) in enumerate(((1, 0), (1, 0), (2, 1), (1, None), (1, 0))):
        b = capi.progression(0, g, 7 + i, 3, n)
        regs.append(zkm2.RegisteredBases("bls12_381", g, b, device=dev, shard=(dev is None), precompute=(i % 2 == 0)))
        outs.append(torch.zeros(12 * g + 1, dtype=torch.int64, device="cuda"))
        want.append(capi.msm(0, g, b, scal))
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        msm_batch_device([(regs[i], d_s.data_ptr(), n, outs[i].data_ptr()) for i in range(5)], stream=st.cuda_stream)
        st.synchronize()
    for i in range(5):
        r = outs[i].cpu().numpy().view(np.uint64)
        assert bool(r[-1]) == want[i][1] and np.array_equal(r[:-1], want[i][0]), i
        regs[i].release()


def test_groth16_create_proof_places_g2_remotely(zkm2):
    """The host-layer prover with a process that owns two devices (g2_device defaults to 1): bytes unchanged."""
    import random
    from oracle.py import groth16_exact as gx, groth16_setup as gs
    from test_groth16_verify import _gpu_proving_key, _prove_exact, C, P   # tests/ is on sys.path (pytest rootdir mode)
    from zkmember_b200.groth16 import create_proof
    cs, z = gs.random_circuit(num_constraints=13, num_inputs=2, seed=5, p=P)
    par = gs.generate_parameters(C, cs, seed=4)
    r, s = random.Random(1).randrange(P), random.Random(2).randrange(P)
    a, b, c = cs.evaluation_vectors(z, P)
    zpk = _gpu_proving_key(par["pk"], True)
    try:
        mont = lambda v: capi.ints_to_limbs([C.fr.to_mont(x) for x in v], C.fr.limbs64)
        canon = lambda v: capi.ints_to_limbs(v, C.fr.limbs64)
        proof = create_proof(zpk, r, s, mont(a), mont(b), mont(c), canon(z[1:cs.num_instance]), canon(z[cs.num_instance:]))
    finally:
        zpk.release()
    assert proof.serialize() == gx.serialize_proof(C, *_prove_exact(cs, z, par, r, s))


def test_host_transforms_spread_over_devices(zkm2):
    from concurrent.futures import ThreadPoolExecutor
    zkm2.set_option("spread_host_calls", 1)
    try:
        x = capi.random_field_elements(0, 1 << 12, seed=8)
        want = capi.ntt(0, x)
        dom = zkm2.Radix2EvaluationDomain("bls12_381", 12)
        with ThreadPoolExecutor(max_workers=4) as ex:
            for got in ex.map(lambda _: dom.fft(x), range(8)):
                assert np.array_equal(got, want)
    finally:
        zkm2.set_option("spread_host_calls", 0)


def test_kzg_open_over_sharded_powers(zkm2):
    from zkmember_b200.kzg import KZG10, Powers
    n = 4000
    powers = capi.progression(0, 1, 5, 3, n)
    coeffs = capi.random_field_elements(0, n, seed=2)
    z = capi.random_field_elements(0, 1, seed=3)[0]
    one, sh = Powers("bls12_381", powers), Powers("bls12_381", powers, shard=True)
    try:
        assert KZG10.open(one, coeffs, z).w == KZG10.open(sh, coeffs, z).w
        assert KZG10.commit(one, coeffs) == KZG10.commit(sh, coeffs)
    finally:
        one.release()
        sh.release()


def test_registration_cache_of_the_literal_msm_call(zkm2):
    """zkm_msm_g1(bases, scalars) keeps the uploaded bases; a changed vector at the same address is re-uploaded."""
    L = zkm2._lib.lib()
    stats = np.zeros(4, dtype=np.uint64)
    sp = ctypes.c_void_p(stats.ctypes.data)
    zkm2._lib.check(L.zkm_msm_cache_clear())
    n = 4096
    bases = capi.progression(0, 1, 3, 5, n)
    s1, s2 = capi.random_scalars(0, n, 1), capi.random_scalars(0, n, 2, "witness")
    zkm2._lib.check(L.zkm_msm_cache_stats(sp))
    h0, m0 = int(stats[0]), int(stats[1])
    for s in (s1, s2, s1):
        _check(zkm2.VariableBaseMSM.multi_scalar_mul(bases, s), *capi.msm(0, 1, bases, s))
    zkm2._lib.check(L.zkm_msm_cache_stats(sp))
    assert int(stats[0]) - h0 == 2 and int(stats[1]) - m0 == 1 and int(stats[2]) >= 1
    # every sampled position is touched by rewriting the whole vector in place: fingerprint mismatch -> fresh upload
    bases[:] = capi.progression(0, 1, 4, 9, n)
    _check(zkm2.VariableBaseMSM.multi_scalar_mul(bases, s1), *capi.msm(0, 1, bases, s1))
    zkm2._lib.check(L.zkm_msm_cache_stats(sp))
    assert int(stats[1]) - m0 == 2
    # full-content fingerprint catches a single changed record anywhere
    zkm2.set_option("msm_cache", 2)
    try:
        _check(zkm2.VariableBaseMSM.multi_scalar_mul(bases, s2), *capi.msm(0, 1, bases, s2))
        bases[1234] = bases[77]
        _check(zkm2.VariableBaseMSM.multi_scalar_mul(bases, s2), *capi.msm(0, 1, bases, s2))
    finally:
        zkm2.set_option("msm_cache", 1)
    zkm2.set_option("msm_cache", 0)
    try:
        _check(zkm2.VariableBaseMSM.multi_scalar_mul(bases, s1), *capi.msm(0, 1, bases, s1))
    finally:
        zkm2.set_option("msm_cache", 1)
    zkm2._lib.check(L.zkm_msm_cache_clear())


def test_profile_counters_are_exact(zkm2):
    import torch
    L = zkm2._lib.lib()
    n = 1 << 14
    bases = capi.progression(0, 1, 3, 5, n)
    scal = capi.random_scalars(0, n, 4)
    scal[:100] = 0                                   # zero scalars contribute no list entries
    reg = zkm2.RegisteredBases("bls12_381", 1, bases)
    d_s = torch.from_numpy(scal.view(np.int64)).cuda()
    d_out = torch.zeros(13, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    zkm2.set_option("profile", 1)
    zkm2.set_option("msm_window_bits", 11)
    zkm2.set_option("msm_affine_levels", 2)
    try:
        reg.msm_device(d_s.data_ptr(), n, d_out.data_ptr())
        cnt = np.zeros(16, dtype=np.uint64)
        zkm2._lib.check(L.zkm_profile_last_msm_counts(ctypes.c_void_p(cnt.ctypes.data)))
        ms = np.zeros(6)
        zkm2._lib.check(L.zkm_profile_last_msm(ctypes.c_void_p(ms.ctypes.data)))
    finally:
        zkm2.set_option("profile", 0)
        zkm2.set_option("msm_window_bits", 0)
        zkm2.set_option("msm_affine_levels", -1)
        reg.release()
    c, W = 11, (255 + 1 + 10) // 11
    assert int(cnt[0]) == n and int(cnt[1]) == W and int(cnt[2]) == c and int(cnt[12]) == 2
    # entries = non-zero signed digits of the non-zero scalars, recomputed on the host
    v = [sum(int(scal[i, j]) << (64 * j) for j in range(4)) for i in range(n)]
    entries = 0
    for s in v:
        carry = 0
        for w in range(W):
            d = ((s >> (c * w)) & ((1 << c) - 1)) + carry
            carry = 0
            if d > (1 << (c - 1)):
                d = (1 << c) - d
                carry = 1
            entries += 1 if d else 0
    assert int(cnt[3]) == entries
    assert 0 < int(cnt[5]) <= int(cnt[4]) <= entries and int(cnt[4]) >= (entries + 1) // 2
    assert int(cnt[13]) > 0 and (ms >= 0).all()
    want_xy, want_inf = capi.msm(0, 1, bases, scal)
    rec = d_out.cpu().numpy().view(np.uint64)
    assert np.array_equal(rec[:-1], want_xy)


def test_sharded_host_msm_with_chunked_scalar_upload(zkm2):
    """Shards of >= 2^22 scalars take the chunked upload of the host entry points (copy stream + one event per chunk, the
    histogram pass behind the copies) on EVERY owning device: 2^23 points sharded two ways == the unsharded MSM == the
    known-discrete-log identity."""
    import torch
    from oracle.py import exact
    curve = BLS12_381
    log_n = 23
    n = 1 << log_n
    a0, d = 0x1234567, 0x89ABCDE
    W = curve.fq.limbs64
    L = zkm2._lib.lib()
    d_bases = torch.empty((n, 2 * W), dtype=torch.int64, device="cuda:0")
    zkm2._lib.check(L.zkm_testgen_progression_device(curve.curve_id, 1, a0, d, n, ctypes.c_void_p(d_bases.data_ptr()),
                                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    scal = capi.random_scalars(curve.curve_id, n, seed=0x5EED0000 + log_n)
    reg_s = zkm2.RegisteredBases.from_device(curve.name, 1, d_bases.data_ptr(), n, shard=True)
    reg_1 = zkm2.RegisteredBases.from_device(curve.name, 1, d_bases.data_ptr(), n)
    try:
        got = [reg_s.msm(scal), reg_s.msm(scal), reg_1.msm(scal)]
    finally:
        reg_s.release()
        reg_1.release()
        del d_bases
        torch.cuda.empty_cache()
    s_int = scal.astype(object)
    s_vals = sum(s_int[:, j] << (64 * j) for j in range(curve.fr.limbs64))
    k = int(np.sum(s_vals * (a0 + np.arange(n, dtype=object) * d)) % curve.fr.modulus)
    G = exact.Group(curve, 1)
    b, _ = exact.point_to_bytes(curve, 1, G.mul(G.gen, k))
    for g in got:
        assert not g.infinity and g.xy.tobytes() == b

#!/usr/bin/env python3
"""Tiny end-to-end pass over every kernel family for compute-sanitizer --tool memcheck
(one tool per gpurun call, smallest sizes): MSM G1/G2 with and without the affine levels and
precomputed multiples, all NTT pass plans, witness map, KZG commit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkmember_b200 as zkm  # noqa: E402
from zkmember_b200.groth16 import witness_map  # noqa: E402
from zkmember_b200.kzg import KZG10, Powers  # noqa: E402
from oracle import capi  # noqa: E402

zkm.init(0)
ok = True
for curve, cid in (("bls12_381", 0), ("bn254", 1)):
    for g, n in ((1, 700), (2, 150)):
        bases = capi.progression(cid, g, 5, 3, n)
        scal = capi.random_scalars(cid, n, seed=g, kind="witness")
        inf = np.zeros(n, dtype=np.uint8)
        inf[3] = 1
        want = capi.msm(cid, g, bases, scal, inf)
        for levels, pre in ((-1, False), (2, False), (-1, True)):
            zkm.set_option("msm_affine_levels", levels)
            if pre:
                zkm.set_option("msm_precompute", 1)
            reg = zkm.RegisteredBases(curve, g, bases, inf)
            zkm.set_option("msm_precompute", 0)
            got = reg.msm(scal)
            reg.release()
            ok &= got.infinity == want[1] and np.array_equal(got.xy, want[0])
        zkm.set_option("msm_affine_levels", -1)
    for log_n, radix in ((2, 12), (5, 12), (10, 12), (13, 6), (13, 12)):
        zkm.set_option("ntt_max_radix_log", radix)
        x = capi.random_field_elements(cid, 1 << log_n, seed=log_n)
        dom = zkm.Radix2EvaluationDomain(curve, log_n)
        ok &= np.array_equal(dom.fft(x), capi.ntt(cid, x))
        ok &= np.array_equal(dom.coset_ifft(x), capi.ntt(cid, x, True, True))
    zkm.set_option("ntt_max_radix_log", 12)
    a, b, c = (capi.random_field_elements(cid, 256, seed=s) for s in (1, 2, 3))
    ok &= np.array_equal(witness_map(a, b, c, curve=curve), capi.witness_map(cid, a, b, c))
    pw = Powers(curve, capi.progression(cid, 1, 9, 2, 300))
    co = capi.random_field_elements(cid, 300, seed=4)
    co[:5] = 0
    fid = 1 if cid == 0 else 3
    rep = np.stack([capi.field_op(fid, 4, co[i]) for i in range(5, 300)])
    w = capi.msm(cid, 1, capi.progression(cid, 1, 9, 2, 300)[5:], rep)
    gk = KZG10.commit(pw, co)
    ok &= gk.infinity == w[1] and np.array_equal(gk.xy, w[0])
    pw.release()
zkm.shutdown()
print("sanitizer target:", "PARITY OK" if ok else "PARITY MISMATCH")
sys.exit(0 if ok else 3)

#!/usr/bin/env python3
"""Tuning sweep of the runtime knobs (pair-level batch sizes, fold fan-in, chunk length) on fixed workloads."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkmember_b200 as zkm
from zkmember_b200 import _lib
from oracle import capi
zkm.init(0); L = _lib.lib()
dev = torch.device("cuda:0"); st = torch.cuda.Stream(); torch.cuda.set_stream(st); sp = ctypes.c_void_p(st.cuda_stream)

def run(reg, d_s, n, d_rec, reps=3):
    f = lambda: reg.msm_device(d_s.data_ptr(), n, d_rec.data_ptr(), stream=st.cuda_stream)
    for _ in range(2): f()
    torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]

def setup(lg, group=1, pre=False, kind="uniform"):
    n = 1 << lg; W = 6 * group
    d_b = torch.empty((n, 2 * W), dtype=torch.int64, device=dev)
    _lib.check(L.zkm_testgen_progression_device(0, group, 0x1234567, 0x89ABCDE, n, ctypes.c_void_p(d_b.data_ptr()), sp))
    torch.cuda.synchronize()
    if pre: zkm.set_option("msm_precompute", 1)
    reg = zkm.RegisteredBases.from_device(0, group, d_b.data_ptr(), n)
    zkm.set_option("msm_precompute", 0)
    d_s = torch.from_numpy(capi.random_scalars(0, n, seed=lg, kind=kind).view(np.int64)).to(dev)
    return reg, d_s, n, torch.zeros(2 * W + 1, dtype=torch.int64, device=dev)

big = setup(24)
for m, m2 in ((32, 32), (64, 32), (128, 32), (64, 64), (128, 64), (256, 32)):
    zkm.set_option("msm_pair_m", m); zkm.set_option("msm_pair_m2", m2)
    print(json.dumps({"n": 24, "pair_m": m, "pair_m2": m2, "ms": run(*big)}), flush=True)
zkm.set_option("msm_pair_m", 64); zkm.set_option("msm_pair_m2", 32)
for lv in (3, 4, 5, 6):
    zkm.set_option("msm_affine_levels", lv)
    print(json.dumps({"n": 24, "affine_levels": lv, "ms": run(*big)}), flush=True)
zkm.set_option("msm_affine_levels", -1)
for c in (19, 20, 21):
    zkm.set_option("msm_window_bits", c)
    print(json.dumps({"n": 24, "window_bits": c, "ms": run(*big)}), flush=True)
zkm.set_option("msm_window_bits", 0)
big[0].release(); del big
for name, args in (("g1_2p16_pre_witness", dict(lg=16, pre=True, kind="witness")), ("g2_2p16_pre_witness", dict(lg=16, group=2, pre=True, kind="witness")),
                   ("g1_2p16_plain", dict(lg=16))):
    w = setup(**args)
    for fold in (4, 8, 16, 32):
        for chunk in (0, 8, 32, 64):
            zkm.set_option("msm_fold", fold); zkm.set_option("msm_chunk", chunk)
            print(json.dumps({"case": name, "fold": fold, "chunk": chunk, "ms": run(*w, reps=5)}), flush=True)
    zkm.set_option("msm_fold", 0); zkm.set_option("msm_chunk", 0)
    w[0].release()

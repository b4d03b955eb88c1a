#!/usr/bin/env python3
"""Window-size sweep for small registered MSMs with precomputed multiples (latency-bound regime)."""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkmember_b200 as zkm
from zkmember_b200 import _lib
from oracle import capi
zkm.init(0); L = _lib.lib()
dev = torch.device("cuda:0"); st = torch.cuda.Stream(); torch.cuda.set_stream(st); sp = ctypes.c_void_p(st.cuda_stream)
for group in (1, 2):
    for lg in (14, 15, 16, 17):
        if group == 2 and lg > 18: continue
        n = 1 << lg; W = 6 * group
        d_b = torch.empty((n, 2 * W), dtype=torch.int64, device=dev)
        _lib.check(L.zkm_testgen_progression_device(0, group, 0x1234567, 0x89ABCDE, n, ctypes.c_void_p(d_b.data_ptr()), sp))
        torch.cuda.synchronize()
        d_s = torch.from_numpy(capi.random_scalars(0, n, seed=lg, kind="witness").view(np.int64)).to(dev)
        d_rec = torch.zeros(2 * W + 1, dtype=torch.int64, device=dev)
        for c in (0, 6, 7, 8, 9, 10):
            zkm.set_option("msm_window_bits", c); zkm.set_option("msm_precompute", 1)
            reg = zkm.RegisteredBases.from_device(0, group, d_b.data_ptr(), n)
            zkm.set_option("msm_precompute", 0)
            f = lambda: reg.msm_device(d_s.data_ptr(), n, d_rec.data_ptr(), stream=st.cuda_stream)
            for _ in range(2): f()
            torch.cuda.synchronize(); ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            print(json.dumps({"group": group, "log_n": lg, "c": c, "ms": sorted(ts)[2]}), flush=True)
            reg.release()
        zkm.set_option("msm_window_bits", 0)

#!/usr/bin/env python3
"""Small fixed workload for ncu captures: one warm + one measured 2^LOG G1 MSM and one warm + one measured
2^LOG Fr NTT, device resident.  Usage: python tools/profile_target.py [log_n] [curve id: 0 bls12_381 | 1 bn254 | 2 bw6_761]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkmember_b200 as zkm  # noqa: E402
from zkmember_b200 import _lib  # noqa: E402
from oracle import capi  # noqa: E402  (input generator only)

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << log_n
cid = int(sys.argv[2]) if len(sys.argv) > 2 else 0
W = capi.coord_words(cid, 1)
zkm.init(0)
L = _lib.lib()
dev = torch.device("cuda:0")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
sp = ctypes.c_void_p(st.cuda_stream)
d_bases = torch.empty((n, 2 * W), dtype=torch.int64, device=dev)
_lib.check(L.zkm_testgen_progression_device(cid, 1, 0x1234567, 0x89ABCDE, n, ctypes.c_void_p(d_bases.data_ptr()), sp))
torch.cuda.synchronize()
reg = zkm.RegisteredBases.from_device(cid, 1, d_bases.data_ptr(), n)
del d_bases
d_scal = torch.from_numpy(capi.random_scalars(cid, n, seed=0x5EED0000 + log_n).view(np.int64)).to(dev)
d_rec = torch.zeros(2 * W + 1, dtype=torch.int64, device=dev)
for _ in range(2):
    reg.msm_device(d_scal.data_ptr(), n, d_rec.data_ptr(), stream=st.cuda_stream)
torch.cuda.synchronize()
x = torch.from_numpy(capi.random_field_elements(cid, n, seed=0x5EED1000 + log_n).view(np.int64)).to(dev)
y = torch.empty_like(x)
for _ in range(2):
    _lib.check(L.zkm_ntt_device(cid, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), log_n, 0, 0, sp))
torch.cuda.synchronize()
print("profile target done", d_rec.cpu().numpy()[:2])

#!/usr/bin/env python3
"""One-GPU check of the sharded HOST path with shards large enough for the chunked scalar upload (>= 2^22 scalars per
shard): zkm_init_devices([0, 0]) -- two lane sets on one GPU --, 2^23 bases registered with ZKM_REG_SHARD, the result
compared with the same MSM over an unsharded registration and with the known-discrete-log identity."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkmember_b200 as zkm
from zkmember_b200 import _lib
from oracle import capi
from oracle.py import exact
from oracle.py.params import BLS12_381 as curve

L = zkm.load()
devs = [0, 1] if L.zkm_device_count() >= 2 else [0, 0]
zkm.init(devs)
log_n = 23
n = 1 << log_n
a0, d = 0x1234567, 0x89ABCDE
W = curve.fq.limbs64
d_bases = torch.empty((n, 2 * W), dtype=torch.int64, device="cuda:0")
_lib.check(L.zkm_testgen_progression_device(0, 1, a0, d, n, ctypes.c_void_p(d_bases.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
torch.cuda.synchronize()
scal = capi.random_scalars(0, n, seed=0x5EED0000 + log_n)
reg_s = zkm.RegisteredBases.from_device("bls12_381", 1, d_bases.data_ptr(), n, shard=True)
reg_1 = zkm.RegisteredBases.from_device("bls12_381", 1, d_bases.data_ptr(), n)
got_s = [reg_s.msm(scal) for _ in range(2)]
got_1 = reg_1.msm(scal)
s_int = scal.astype(object)
s_vals = sum(s_int[:, j] << (64 * j) for j in range(curve.fr.limbs64))
k = int(np.sum(s_vals * (a0 + np.arange(n, dtype=object) * d)) % curve.fr.modulus)
G = exact.Group(curve, 1)
b, f = exact.point_to_bytes(curve, 1, G.mul(G.gen, k))
ok = all((not g.infinity) and g.xy.tobytes() == b for g in got_s + [got_1])
print("devices", devs, "sharded host MSM 2^%d with chunked upload:" % log_n, "OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)

#!/usr/bin/env python3
"""Debug aid: hierarchical window reduction at small n (forced window bits) against the CPU oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkmember_b200 as zkm
from oracle import capi
zkm.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 15
cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [12, 13, 14, 15, 16]
bases = capi.progression(0, 1, 5, 3, n)
scal = capi.random_scalars(0, n, seed=3)
want = capi.msm(0, 1, bases, scal)
reg = zkm.RegisteredBases("bls12_381", 1, bases)
for c in cs:
    zkm.set_option("msm_window_bits", c)
    got = reg.msm(scal)
    print("n", n, "c", c, "ok", got.infinity == want[1] and np.array_equal(got.xy, want[0]), flush=True)

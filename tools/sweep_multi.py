#!/usr/bin/env python3
"""BASELINE.json configs[1] / [2] at 1/2/4/8 GPUs from ONE process (the deployment of the Rust prover: zkm_init_mask).

  python tools/sweep_multi.py --gpus N [--curve bls12_381] [--sizes 20,22,24,26]

MSM: the bases of every size are registered with ZKM_REG_SHARD (range sharding over the N GPUs); one step =
zkm_msm_registered with ALL scalars in pinned host memory (upload of every shard's slice, N pipelines, NVLink P2P gather of
the N result records, k_points_sum on device 0, read-back) -- wall clock, median of --reps; the result is checked by the
known-discrete-log identity.  NTT: N independent 2^k transforms issued from N host threads through zkm_ntt (pinned host
buffers, `spread_host_calls`), aggregate transforms/s; every output is compared with the first.  One JSON line per point."""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkmember_b200 as zkm  # noqa: E402
from zkmember_b200 import _lib  # noqa: E402
from oracle import capi, checks  # noqa: E402  (input generator + result checker, outside the timed regions)

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--curve", default="bls12_381")
ap.add_argument("--sizes", default="20,22,24")
ap.add_argument("--ntt-sizes", default="20,24")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
cid = {"bls12_381": 0, "bn254": 1, "bw6_761": 2}[args.curve]
N = args.gpus
zkm.init(list(range(N)) if N > 1 else 0)
L = _lib.lib()
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
W = capi.coord_words(cid, 1)
a0, d = 0x1234567, 0x89ABCDE
for lg in [int(x) for x in args.sizes.split(",") if x]:
    n = 1 << lg
    d_b = torch.empty((n, 2 * W), dtype=torch.int64, device=dev)
    _lib.check(L.zkm_testgen_progression_device(cid, 1, a0, d, n, ctypes.c_void_p(d_b.data_ptr()), ctypes.c_void_p(0)))
    torch.cuda.synchronize()
    reg = zkm.RegisteredBases.from_device(cid, 1, d_b.data_ptr(), n, shard=(N > 1))
    del d_b
    torch.cuda.empty_cache()
    scal = capi.random_scalars(cid, n, seed=0x5EED0000 + lg)
    h_s = torch.from_numpy(scal.view(np.int64)).pin_memory()
    out = np.zeros(2 * W, dtype=np.uint64)
    inf = np.zeros(1, dtype=np.uint8)
    ts = []
    for it in range(2 + args.reps):
        t0 = time.perf_counter()
        _lib.check(L.zkm_msm_registered(reg.handle, 0, ctypes.c_void_p(h_s.data_ptr()), n, ctypes.c_void_p(out.ctypes.data),
                                        ctypes.c_void_p(inf.ctypes.data)))
        if it >= 2:
            ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    rec = np.concatenate([out, np.array([int(inf[0])], dtype=np.uint64)])
    ok = checks.msm_identity_ok(cid, 1, rec, checks.dlog_sum(scal, a0, d, capi.CURVES[cid].fr.modulus))
    print(json.dumps({"op": "msm_single_process", "curve": args.curve, "gpus": N, "log_n": lg, "e2e_ms": ts[len(ts) // 2],
                      "e2e_ms_best": ts[0], "h2d_bytes": int(n * 8 * capi.fr_words(cid)), "check": bool(ok),
                      "check_kind": "known-discrete-log identity (exact)"}), flush=True)
    reg.release()
    del h_s

zkm.set_option("spread_host_calls", 1)
S = capi.fr_words(cid)
for lg in [int(x) for x in args.ntt_sizes.split(",") if x]:
    n = 1 << lg
    x = capi.random_field_elements(cid, n, seed=0x5EED1000 + lg)
    bufs = [torch.from_numpy(x.view(np.int64).copy()).pin_memory() for _ in range(N)]
    rounds = 4

    def work(b):
        for r in range(rounds):
            _lib.check(L.zkm_ntt(cid, ctypes.c_void_p(b.data_ptr()), lg, 0, 0))      # in place: fft of the previous output
    chk = torch.from_numpy(x.view(np.int64).copy()).pin_memory()
    for _ in range(N):                              # warm: tables on every device come with first use
        _lib.check(L.zkm_ntt(cid, ctypes.c_void_p(chk.data_ptr()), lg, 0, 0))
        ok1 = lg > 24 or np.array_equal(chk.numpy().view(np.uint64).reshape(-1, S), capi.ntt(cid, x))
        chk.copy_(torch.from_numpy(x.view(np.int64)))
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(b,)) for b in bufs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    wall = time.perf_counter() - t0
    ok = bool(ok1) and all(torch.equal(b, bufs[0]) for b in bufs)
    print(json.dumps({"op": "ntt_single_process", "curve": args.curve, "gpus": N, "log_n": lg, "transforms": N * rounds,
                      "wall_ms_per_transform": wall * 1e3 / (N * rounds), "transforms_per_s": N * rounds / wall,
                      "note": "zkm_ntt in place on pinned host buffers from N host threads, spread over the devices (PCIe-bound: "
                              "upload + download per transform)", "check": bool(ok)}), flush=True)
zkm.shutdown()

#!/usr/bin/env python3
"""Standalone sweeps (BASELINE.json configs 2 and 3): G1/G2 MSM and Fr NTT over a range of sizes,
device-resident timings with CUDA events, one JSON line per point.

  python tools/sweep.py msm --curve bls12_381 --group 1 --min 14 --max 24 [--kind uniform|witness]
  python tools/sweep.py ntt --curve bls12_381 --min 14 --max 26
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkmember_b200 as zkm  # noqa: E402
from zkmember_b200 import _lib  # noqa: E402
from oracle import capi, checks  # noqa: E402  (input generator + result checker only, outside the timed regions)

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["msm", "ntt"])
ap.add_argument("--curve", default="bls12_381")
ap.add_argument("--group", type=int, default=1)
ap.add_argument("--min", type=int, default=14)
ap.add_argument("--max", type=int, default=24)
ap.add_argument("--kind", default="uniform")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--precompute", action="store_true")
ap.add_argument("--opt", action="append", default=[], help="library option key=value (zkm_set_option), repeatable")
ap.add_argument("--cpu", action="store_true", help="also time the CPU oracle (arkworks algorithm restated) up to 2^20")
args = ap.parse_args()

cid = {"bls12_381": 0, "bn254": 1, "bw6_761": 2}[args.curve]
zkm.init(0)
for kv in args.opt:
    k_, v_ = kv.split("=")
    zkm.set_option(k_, int(v_))
L = _lib.lib()
dev = torch.device("cuda:0")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
sp = ctypes.c_void_p(st.cuda_stream)


def timeit(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


if args.what == "msm":
    W = capi.coord_words(cid, args.group)
    nmax = 1 << args.max
    d_bases = torch.empty((nmax, 2 * W), dtype=torch.int64, device=dev)
    _lib.check(L.zkm_testgen_progression_device(cid, args.group, 0x1234567, 0x89ABCDE, nmax,
                                                ctypes.c_void_p(d_bases.data_ptr()), sp))
    torch.cuda.synchronize()
    del_after = True
    d_rec = torch.zeros(2 * W + 1, dtype=torch.int64, device=dev)
    for lg in range(args.min, args.max + 1):
        n = 1 << lg
        reg = zkm.RegisteredBases.from_device(cid, args.group, d_bases.data_ptr(), n, precompute=args.precompute)   # first n bases
        h = capi.random_scalars(cid, n, seed=0x5EED0000 + lg, kind=args.kind)
        d_s = torch.from_numpy(h.view(np.int64)).to(dev)
        zkm.set_option("profile", 1)
        med, best = timeit(lambda: reg.msm_device(d_s.data_ptr(), n, d_rec.data_ptr(), stream=st.cuda_stream), args.reps)
        stages = np.zeros(6)
        _lib.check(L.zkm_profile_last_msm(ctypes.c_void_p(stages.ctypes.data)))
        row = {"op": "msm", "curve": args.curve, "group": args.group, "log_n": lg, "kind": args.kind, "ms": med,
               "ms_best": best, "window_bits": zkm.msm_window_bits(cid, args.group, n),
               "stage_ms": dict(zip(["sort", "affine", "tasks", "accumulate", "fold", "reduce"], [round(float(v), 4) for v in stages]))}
        if args.cpu and lg <= 20:
            import time
            hb = capi.progression(cid, args.group, 0x1234567, 0x89ABCDE, n)
            t0 = time.perf_counter()
            capi.msm(cid, args.group, hb, h)
            row["cpu_ms"] = (time.perf_counter() - t0) * 1e3
        # result check of what was timed: known-discrete-log identity, exact big-int arithmetic (oracle/checks.py)
        rec = d_rec.cpu().numpy().view(np.uint64)
        k_sum = checks.dlog_sum(h, 0x1234567, 0x89ABCDE, capi.CURVES[cid].fr.modulus)
        row["check"] = checks.msm_identity_ok(cid, args.group, rec, k_sum)
        row["check_kind"] = "known-discrete-log identity (exact)"
        row["precompute"] = bool(args.precompute)
        if args.opt:
            row["opts"] = args.opt
        reg.release()
        print(json.dumps(row), flush=True)
else:
    for lg in range(args.min, args.max + 1):
        n = 1 << lg
        hx = capi.random_field_elements(cid, n, seed=0x5EED1000 + lg)
        x = torch.from_numpy(hx.view(np.int64)).to(dev)
        y = torch.empty_like(x)
        row = {"op": "ntt", "curve": args.curve, "log_n": lg}
        for name, inv, cos in (("fft", 0, 0), ("ifft", 1, 0), ("coset_fft", 0, 1), ("coset_ifft", 1, 1)):
            med, best = timeit(lambda: _lib.check(L.zkm_ntt_device(cid, ctypes.c_void_p(x.data_ptr()),
                                                                   ctypes.c_void_p(y.data_ptr()), lg, inv, cos, sp)), args.reps)
            row[name + "_ms"] = med
        # result check: every byte against the CPU oracle up to 2^24; above that round trip + Horner spot checks
        _lib.check(L.zkm_ntt_device(cid, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()), lg, 0, 0, sp))
        torch.cuda.synchronize()
        if lg <= 24:
            row["check"] = bool(np.array_equal(y.cpu().numpy().view(np.uint64), capi.ntt(cid, hx)))
            row["check_kind"] = "fft bytes == CPU oracle"
        else:
            z = torch.empty_like(x)
            _lib.check(L.zkm_ntt_device(cid, ctypes.c_void_p(y.data_ptr()), ctypes.c_void_p(z.data_ptr()), lg, 1, 0, sp))
            torch.cuda.synchronize()
            ok = bool(torch.equal(z, x))
            z.zero_()
            z[:1024] = x[:1024]
            _lib.check(L.zkm_ntt_device(cid, ctypes.c_void_p(z.data_ptr()), ctypes.c_void_p(y.data_ptr()), lg, 0, 0, sp))
            torch.cuda.synchronize()
            ev = {k: y[k].cpu().numpy().view(np.uint64) for k in (1, 4097, n - 1)}
            row["check"] = ok and checks.horner_ok(cid, lg, hx[:1024], ev)
            row["check_kind"] = "ifft(fft(x)) == x and Horner spot checks (exact)"
            del z
        row["hbm_frac_fft"] = 2.0 * 8 * capi.fr_words(cid) * n / (row["fft_ms"] * 1e-3) / 6539.9e9
        if args.cpu and lg <= 22:
            import time
            t0 = time.perf_counter()
            capi.ntt(cid, hx)
            row["cpu_fft_ms"] = (time.perf_counter() - t0) * 1e3
        print(json.dumps(row), flush=True)
        del x, y

#!/usr/bin/env python3
"""zkMember-shaped Marlin proof PROXY (BASELINE.json config 4): the KZG10 commitments and radix-2 transforms of
one `Marlin::prove` (/root/reference/benches/marlin.rs:202,311 -> ark-marlin 0.3.0 AHP prover rounds 1-3 and
MarlinKZG10::commit / open in ark-poly-commit 0.3.0) for a constraint system with |H| = 2^log_h constraints /
variables and |K| = 2^log_k non-zero matrix entries.

Work per proof (sizes from the degree bounds of the AHP):

  round 1   HIDING commits of w, z_a, z_b (|H| coefficients each) and the mask polynomial (3|H|)   3 iffts of size |H|
  round 2   commit t, h_1 (2|H|) and HIDING commit of g_1 (|H|)                                     5 ffts + 1 ifft on the 4|H| domain
  round 3   commit g_2 (|K|) and h_2 (6|K|)                                                        1 ifft |K|, 4 ffts + 1 ifft on 4|K|
  opening   two KZG10::open calls with hiding: polynomials of 3|H| and 6|K| coefficients divided by (X - z) ON THE DEVICE,
            witness commitments over powers_of_g, hiding witnesses over powers_of_gamma_g, random_v

i.e. nine commitments (five with their hiding MSM over powers_of_gamma_g) + two openings over slices of one registered SRS
(`powers_of_g` / `powers_of_gamma_g`, registered once like the committer key is built once in the bench).  The non-hiding
commitments of a round go through ONE `zkm_kzg_commit_batch` call, the hiding ones and the openings are issued concurrently
from host threads, rounds in sequence (Fiat-Shamir).  Coefficients are uniform Fr elements in Montgomery form handed in as
HOST arrays; the library does the leading-zero skip, into_repr, the division and the MSMs on the device.  It is a proxy: constraint
synthesis, the sumcheck polynomial arithmetic and Fiat-Shamir hashing are host Rust and not included, and the
transform list approximates the AHP's volume (shape, not a transcript).  Prints one JSON line.

  python tools/marlin_proxy.py [--log-h 16] [--log-k 18] [--proofs 6] [--curve bls12_381|bw6_761] [--precompute] [--cpu]
"""
import argparse
import ctypes
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import zkmember_b200 as zkm  # noqa: E402
from zkmember_b200 import _lib  # noqa: E402
from zkmember_b200.kzg import KZG10  # noqa: E402
from oracle import capi  # noqa: E402  (input generator / CPU baseline / checker only)

ap = argparse.ArgumentParser()
ap.add_argument("--log-h", type=int, default=16)
ap.add_argument("--log-k", type=int, default=18)
ap.add_argument("--proofs", type=int, default=6)
ap.add_argument("--curve", default="bls12_381")
ap.add_argument("--precompute", action="store_true", help="register the SRS with precomputed window multiples (static committer key)")
ap.add_argument("--cpu", action="store_true", help="check every commitment against / time the CPU restatement (1 proof)")
ap.add_argument("--device", type=int, default=int(os.environ.get("LOCAL_RANK", "0")))
args = ap.parse_args()

cid = {"bls12_381": 0, "bn254": 1, "bw6_761": 2}[args.curve]
H, K = 1 << args.log_h, 1 << args.log_k
W1, SW = capi.coord_words(cid, 1), capi.fr_words(cid)
torch.cuda.set_device(args.device)
zkm.init(args.device)
L = _lib.lib()
dev = torch.device("cuda", args.device)
st = torch.cuda.Stream(device=dev)
sp = ctypes.c_void_p(st.cuda_stream)

# ---- the rounds: (name, coefficients) per commit, (log size, inverse, coset) per transform
HIDING = {"w", "z_a", "z_b", "mask", "g_1"}
ROUNDS = [
    {"commits": [("w", H), ("z_a", H), ("z_b", H), ("mask", 3 * H)], "opens": [],
     "ntts": [(args.log_h, 1, 0)] * 3},
    {"commits": [("t", H), ("g_1", H), ("h_1", 2 * H)], "opens": [],
     "ntts": [(args.log_h + 2, 0, 1)] * 5 + [(args.log_h + 2, 1, 1)]},
    {"commits": [("g_2", K), ("h_2", 6 * K)], "opens": [],
     "ntts": [(args.log_k, 1, 0)] + [(args.log_k + 2, 0, 1)] * 4 + [(args.log_k + 2, 1, 1)]},
    {"commits": [], "opens": [("open_beta", 3 * H), ("open_gamma", 6 * K)], "ntts": []},
]
n_srs = max(6 * K, 3 * H)

# ---- committer key: powers_of_g registered once (any G1 points serve for timing and parity)
t0 = time.perf_counter()
d_srs = torch.empty((n_srs, 2 * W1), dtype=torch.int64, device=dev)
_lib.check(L.zkm_testgen_progression_device(cid, 1, 0x51D5, 0x7, n_srs, ctypes.c_void_p(d_srs.data_ptr()), sp))
torch.cuda.synchronize()
powers = zkm.RegisteredBases.from_device(cid, 1, d_srs.data_ptr(), n_srs, precompute=args.precompute)
reg_s = time.perf_counter() - t0

# ---- polynomials (host, Montgomery Fr) and transform buffers (device)
polys = {}
for r in ROUNDS:
    for i, (name, m) in enumerate(r["commits"] + r["opens"]):
        polys[name] = capi.random_field_elements(cid, m, seed=0xA11CE + len(polys))
# powers_of_gamma_g (hiding bound 2: three powers) and one blinding polynomial per hiding commitment / opening
gamma_host = capi.progression(cid, 1, 0x6A33A, 0x5, 4)
gamma = zkm.RegisteredBases(cid, 1, gamma_host, precompute=args.precompute)
blind = {name: capi.random_field_elements(cid, 3, seed=0xB100 + i) for i, name in enumerate(sorted(polys))}
points = {name: capi.random_field_elements(cid, 1, seed=0xE7A + i)[0] for i, name in enumerate(sorted(polys))}
max_log = max([lg for r in ROUNDS for (lg, _, _) in r["ntts"]] + [1])
d_x = torch.from_numpy(capi.random_field_elements(cid, 1 << max_log, seed=99).view(np.int64)).to(dev)
d_y = torch.empty_like(d_x)
pool = ThreadPoolExecutor(max_workers=6)


def prove_once():
    out = {}
    for r in ROUNDS:
        for (lg, inv, cos) in r["ntts"]:
            _lib.check(L.zkm_ntt_device(cid, ctypes.c_void_p(d_x.data_ptr()), ctypes.c_void_p(d_y.data_ptr()), lg, inv, cos, sp))
        plain = [name for (name, _) in r["commits"] if name not in HIDING]
        futs = [(name, pool.submit(KZG10.commit, powers, polys[name], gamma, blind[name])) for (name, _) in r["commits"] if name in HIDING]
        futs += [(name, pool.submit(KZG10.open, powers, polys[name], points[name], gamma, blind[name])) for (name, _) in r["opens"]]
        fb = pool.submit(KZG10.commit_batch, powers, [polys[n_] for n_ in plain]) if plain else None
        st.synchronize()                      # the transforms of the round
        for name, f in futs:
            out[name] = f.result()            # the round's commitments feed Fiat-Shamir before the next round
        if fb:
            for name, pt in zip(plain, fb.result()):
                out[name] = pt
    return out


for _ in range(2):
    res = prove_once()
torch.cuda.synchronize()
_lib.launch_count(reset=True)
t0 = time.perf_counter()
for _ in range(args.proofs):
    res = prove_once()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / args.proofs
launches = _lib.launch_count() // args.proofs

msm_points = sum(m for r in ROUNDS for (_, m) in r["commits"] + r["opens"])
ntt_elems = sum(1 << lg for r in ROUNDS for (lg, _, _) in r["ntts"])
out = {"op": "marlin_proxy", "curve": args.curve, "log_h": args.log_h, "log_k": args.log_k, "proofs_timed": args.proofs,
       "ms_per_proof": wall * 1e3, "proofs_per_s": 1.0 / wall, "commits_per_proof": sum(len(r["commits"]) for r in ROUNDS),
       "hiding_commits_per_proof": len(HIDING), "openings_per_proof": sum(len(r["opens"]) for r in ROUNDS),
       "msm_points_per_proof": int(msm_points), "largest_msm": int(max(m for r in ROUNDS for (_, m) in r["commits"] + r["opens"])),
       "ntts_per_proof": sum(len(r["ntts"]) for r in ROUNDS), "ntt_elements_per_proof": int(ntt_elems),
       "precompute": bool(args.precompute), "kernel_launches_per_proof": int(launches), "srs_points": int(n_srs), "srs_register_s": reg_s,
       "h2d_bytes_per_proof": int(msm_points * 8 * SW), "d2h_bytes_per_proof": int(11 * (2 * W1 * 8 + 1) + 2 * 8 * SW),
       "note": "KZG10 commits (hiding where the AHP hides) + openings (device division, hiding witness) + radix-2 transforms of "
               "Marlin::prove; no synthesis / sumcheck arithmetic / Fiat-Shamir; transform list approximates the AHP's volume"}

if args.cpu:
    host_srs = d_srs.cpu().numpy().view(np.uint64)
    ok = True
    t_cpu_ntt = 0.0
    hx = d_x.cpu().numpy().view(np.uint64)
    for r in ROUNDS:
        for (lg, inv, cos) in r["ntts"]:
            seg = hx[: 1 << lg]
            t1 = time.perf_counter()
            capi.ntt(cid, seg, bool(inv), bool(cos))
            t_cpu_ntt += time.perf_counter() - t1
    from oracle.py import exact, kzg_exact as kx
    from oracle.py.params import CURVES_BY_ID
    curve = CURVES_BY_ID[cid]
    fr = curve.fr
    G = exact.Group(curve, 1)

    def padd(p1, p2):                        # sum of two (xy limbs, infinity) points with the exact group law
        P1 = None if p1[1] else exact.point_from_bytes(curve, 1, np.asarray(p1[0], dtype=np.uint64).tobytes(), 0)
        P2 = None if p2[1] else exact.point_from_bytes(curve, 1, np.asarray(p2[0], dtype=np.uint64).tobytes(), 0)
        b, f = exact.point_to_bytes(curve, 1, G.add(P1, P2))
        return np.frombuffer(b, dtype=np.uint64), bool(f)

    t_cpu_msm = 0.0
    for r in ROUNDS:
        for (name, m) in r["commits"]:
            t1 = time.perf_counter()
            sc = capi.fr_into_repr(cid, polys[name])
            xy, isinf = capi.msm(cid, 1, host_srs[:m], sc)
            if name in HIDING:
                hxy, hinf = capi.msm(cid, 1, gamma_host[:3], capi.fr_into_repr(cid, blind[name]))
                xy, isinf = padd((xy, isinf), (hxy, hinf))
            t_cpu_msm += time.perf_counter() - t1
            got = res[name]
            ok = ok and got.infinity == isinf and np.array_equal(got.xy, xy)
        for (name, m) in r["opens"]:
            # the division in exact big-int arithmetic (not timed: O(n) next to the MSM), the MSMs by the C++ restatement
            z = fr.from_mont(capi.limbs_to_ints(points[name][None, :])[0])
            ci = [fr.from_mont(v) for v in capi.limbs_to_ints(polys[name])]
            bi = [fr.from_mont(v) for v in capi.limbs_to_ints(blind[name])]
            wq = capi.ints_to_limbs(kx.witness_polynomial(fr.modulus, ci, z), fr.limbs64)
            hq = capi.ints_to_limbs(kx.witness_polynomial(fr.modulus, bi, z), fr.limbs64)
            t1 = time.perf_counter()
            xy, isinf = capi.msm(cid, 1, host_srs[:m - 1], wq)
            hxy, hinf = capi.msm(cid, 1, gamma_host[:2], hq)
            t_cpu_msm += time.perf_counter() - t1
            xy, isinf = padd((xy, isinf), (hxy, hinf))
            got = res[name]
            ok = ok and got.w.infinity == isinf and np.array_equal(got.w.xy, xy)
            ok = ok and capi.limbs_to_ints(got.random_v[None, :])[0] == fr.to_mont(kx.evaluate(fr.modulus, bi, z))
    out["parity_ok"] = bool(ok)
    out["cpu_ms_per_proof"] = (t_cpu_ntt + t_cpu_msm) * 1e3
    out["cpu_proofs_per_s"] = 1.0 / (t_cpu_ntt + t_cpu_msm)
    out["cpu_msm_ms"] = t_cpu_msm * 1e3
    out["cpu_ntt_ms"] = t_cpu_ntt * 1e3
    out["cpu_threads"] = int(capi.lib().orc_num_threads())
    out["cpu_kind"] = "arkworks-0.3.0 algorithms restated in C++ (oracle/cpp), same inputs"
powers.release()
gamma.release()
print(json.dumps(out))

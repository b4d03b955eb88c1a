"""The C++ host layer (include/zkm_b200.hpp: the compiled-language mirror of the arkworks interfaces above
the C ABI).  CPU: it compiles, links against libzkm_b200.so and refuses to run without a GPU (no CPU
fallback).  GPU: it reproduces the golden vectors and oracle-made witness-map vectors byte for byte."""
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "cpp", "test_host_cpp.cpp")
OUT_DIR = os.path.join(HERE, "cpp", "_build")
EXE = os.path.join(OUT_DIR, "test_host_cpp")
LIB_DIR = os.path.join(ROOT, "zkmember_b200", "lib")


def _build():
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "include", "zkm_b200.hpp"), os.path.join(ROOT, "include", "zkm_b200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-o", EXE, SRC, "-L" + LIB_DIR, "-lzkm_b200",
                               "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + LIB_DIR, "-Wl,-rpath,/usr/local/cuda/lib64"])
    return EXE


def _vectors(path):
    import numpy as np
    from oracle import capi
    from oracle.py import exact
    lines = []
    for v in json.load(open(os.path.join(HERE, "golden", "ntt_vectors.json")))["vectors"]:
        lines.append("ntt %s %d %d %d %s %s" % (v["curve"], v["log_n"], int(v["inverse"]), int(v["coset"]), v["input"], v["output"]))
    for v in json.load(open(os.path.join(HERE, "golden", "msm_vectors.json")))["vectors"]:
        inf = bytes(v["infinity"]).hex() or "-"
        lines.append("msm %s %d %d %s %s %s %s %d" % (v["curve"], v["group"], v["n"], v["bases"] or "-", inf, v["scalars"] or "-",
                                                         v["result"], v["result_infinity"]))
    for cid, name in ((0, "bls12_381"), (1, "bn254"), (2, "bw6_761")):
        n = 1 << 9
        a, b, c = (capi.random_field_elements(cid, n, seed=s) for s in (21, 22, 23))
        h = capi.witness_map(cid, a, b, c)
        lines.append("wmap %s 9 %s %s %s %s" % (name, a.tobytes().hex(), b.tobytes().hex(), c.tobytes().hex(), h.tobytes().hex()))
    # KZG10 commit / hiding commit / open / hiding open: expected points from the C++ oracle's MSM over the exact quotient
    from oracle.py import kzg_exact as kx
    for cid, name in ((0, "bls12_381"), (1, "bn254"), (2, "bw6_761")):
        fr = capi.CURVES[cid].fr
        n = 300
        pw, gpw = capi.progression(cid, 1, 5, 3, n), capi.progression(cid, 1, 11, 7, n)
        co, bl = capi.random_field_elements(cid, n, seed=31), capi.random_field_elements(cid, 3, seed=32)
        zm = capi.random_field_elements(cid, 1, seed=33)
        ci, bi = ([fr.from_mont(v) for v in capi.limbs_to_ints(a)] for a in (co, bl))
        zi = fr.from_mont(capi.limbs_to_ints(zm)[0])
        lim = lambda vals: capi.ints_to_limbs(vals, fr.limbs64)
        group = exact.Group(capi.CURVES[cid], 1)

        def add(p1, p2):
            P1 = None if p1[1] else exact.point_from_bytes(capi.CURVES[cid], 1, p1[0].tobytes(), 0)
            P2 = None if p2[1] else exact.point_from_bytes(capi.CURVES[cid], 1, p2[0].tobytes(), 0)
            b, f = exact.point_to_bytes(capi.CURVES[cid], 1, group.add(P1, P2))
            return np.frombuffer(b, dtype=np.uint64), bool(f)
        c0 = capi.msm(cid, 1, pw, lim(ci))
        r0 = capi.msm(cid, 1, gpw[:3], lim(bi))
        w0 = capi.msm(cid, 1, pw[:n - 1], lim(kx.witness_polynomial(fr.modulus, ci, zi)))
        h0 = capi.msm(cid, 1, gpw[:2], lim(kx.witness_polynomial(fr.modulus, bi, zi)))
        pts = [c0, add(c0, r0), w0, add(w0, h0)]
        rv = lim([fr.to_mont(kx.evaluate(fr.modulus, bi, zi))])
        lines.append("kzg %s %d %s %s %s %s %s %s %s" % (
            name, n, pw.tobytes().hex(), gpw.tobytes().hex(), co.tobytes().hex(), bl.tobytes().hex(), zm.tobytes().hex(),
            " ".join("%s %d" % (np.asarray(p[0], dtype=np.uint64).tobytes().hex(), int(p[1])) for p in pts), rv.tobytes().hex()))
    # arkworks compressed serialization (host-only lines, checked before zkm::init)
    from oracle.py import groth16_exact as gx
    from oracle.py.params import CURVES
    n_ser = 0
    for name in ("bls12_381", "bn254", "bw6_761"):
        curve = CURVES[name]
        for g in (1, 2):
            G = exact.Group(curve, g)
            pts = G.progression(7, 5, 8) + [None]
            pts += [G.neg(P) for P in pts[:4]]
            for P in pts:
                b, f = exact.point_to_bytes(curve, g, P)
                lines.append("ser %s %d %s %d %s" % (name, g, b.hex(), f, gx.serialize_affine(curve, g, P).hex()))
                n_ser += 1
    open(path, "w").write("\n".join(lines) + "\n")
    return len(lines) - n_ser, n_ser


def test_cpp_host_layer_builds_and_has_no_cpu_fallback(tmp_path):
    import torch
    exe = _build()
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    vec = tmp_path / "v.txt"
    _, n_ser = _vectors(str(vec))
    r = subprocess.run([exe, str(vec)], capture_output=True, text=True)
    # the host-only part (arkworks compressed serialization) runs without a GPU and must match the restatement
    assert "ser: %d vectors, 0 mismatches" % n_ser in r.stdout, r.stdout[-2000:]
    assert r.returncode == 3 and "INIT FAILED" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cpp_host_layer_matches_golden_vectors(tmp_path):
    exe = _build()
    vec = tmp_path / "v.txt"
    count, n_ser = _vectors(str(vec))
    r = subprocess.run([exe, str(vec)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ser: %d vectors, 0 mismatches" % n_ser in r.stdout
    assert "host_cpp: %d vectors, 0 mismatches" % count in r.stdout

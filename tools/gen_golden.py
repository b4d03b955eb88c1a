#!/usr/bin/env python3
"""Generate tests/golden/*.json -- small known-answer vectors for the hot path.

The reference (/root/reference) holds no golden vector for MSM / FFT and arkworks
cannot be built here, so these vectors come from the algorithm-independent DEFINITIONS evaluated
with exact Python big integers (oracle/py/exact.py: naive double-and-add MSM, O(n^2) DFT) in
arkworks' byte formats (Montgomery little-endian limbs).  Both the C++ restatement of the
arkworks algorithms (oracle/cpp) and the CUDA library are checked against them.
Run: python tools/gen_golden.py    (deterministic; the output is committed)
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.py import exact  # noqa: E402
from oracle.py.params import BLS12_381, BN254, BW6_761  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def hexs(b: bytes) -> str:
    return b.hex()


def ntt_vectors():
    vecs = []
    for curve in (BLS12_381, BN254, BW6_761):
        fr = curve.fr
        rng = random.Random(1000 + curve.curve_id)
        for log_n in (0, 1, 3, 4, 6):
            n = 1 << log_n
            x = [rng.randrange(fr.modulus) for _ in range(n)]
            if log_n == 3:
                x[0], x[1], x[2] = 0, 1, fr.modulus - 1
            for inverse in (False, True):
                for coset in (False, True):
                    y = exact.ntt_def(fr, x, inverse, coset)
                    vecs.append({
                        "curve": curve.name, "log_n": log_n, "inverse": inverse, "coset": coset,
                        "input": hexs(b"".join(exact.fe_to_bytes(fr, v) for v in x)),
                        "output": hexs(b"".join(exact.fe_to_bytes(fr, v) for v in y)),
                    })
    return vecs


def msm_vectors():
    vecs = []
    for curve in (BLS12_381, BN254, BW6_761):
        for g in (1, 2):
            G = exact.Group(curve, g)
            rng = random.Random(2000 + 10 * curve.curve_id + g)
            for n, kind in ((0, "uniform"), (1, "uniform"), (12 if g == 1 else 6, "mixed")):
                pts = G.progression(rng.randrange(1, 1 << 30), rng.randrange(1, 1 << 30), n)
                scal = [rng.randrange(curve.fr.modulus) for _ in range(n)]
                if kind == "mixed":
                    scal[0] = 0
                    scal[1] = 1
                    scal[2] = curve.fr.modulus - 1
                    pts[3] = None                       # point at infinity
                    pts[5] = pts[4]                     # repeated point
                    if n > 8:
                        pts[7] = G.neg(pts[6])          # P, -P with equal scalars
                        scal[7] = scal[6]
                res = G.msm_naive(pts, scal)
                bb, ff = b"", []
                for P in pts:
                    b, f = exact.point_to_bytes(curve, g, P)
                    bb += b
                    ff.append(f)
                rb, rf = exact.point_to_bytes(curve, g, res)
                vecs.append({
                    "curve": curve.name, "group": g, "n": n,
                    "bases": hexs(bb), "infinity": ff,
                    "scalars": hexs(b"".join(exact.scalar_to_bytes(curve.fr, s) for s in scal)),
                    "result": hexs(rb), "result_infinity": rf,
                })
    return vecs


def wmap_vectors():
    """R1CStoQAP::witness_map (ark-groth16 0.3.0 src/r1cs_to_qap.rs) from the exact O(n^2) definitions, n = 8."""
    from oracle.py import groth16_exact as gx
    vecs = []
    for curve in (BLS12_381, BN254, BW6_761):
        fr = curve.fr
        rng = random.Random(3000 + curve.curve_id)
        a, b, c = ([rng.randrange(fr.modulus) for _ in range(8)] for _ in range(3))
        h = gx.witness_map(curve, a, b, c)
        enc = lambda v: hexs(b"".join(exact.fe_to_bytes(fr, x) for x in v))
        vecs.append({"curve": curve.name, "log_n": 3, "a": enc(a), "b": enc(b), "c": enc(c), "h": enc(h)})
    return vecs


def main():
    os.makedirs(OUT, exist_ok=True)
    json.dump({"generator": "tools/gen_golden.py", "format": "Montgomery LE limbs, hex", "vectors": wmap_vectors()},
              open(os.path.join(OUT, "wmap_vectors.json"), "w"), indent=0)
    json.dump({"generator": "tools/gen_golden.py", "format": "Montgomery LE limbs, hex", "vectors": ntt_vectors()},
              open(os.path.join(OUT, "ntt_vectors.json"), "w"), indent=0)
    json.dump({"generator": "tools/gen_golden.py", "format": "affine Montgomery LE limbs / canonical LE scalars, hex",
               "vectors": msm_vectors()}, open(os.path.join(OUT, "msm_vectors.json"), "w"), indent=0)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

#!/bin/bash
# GPU session r2o: pair kernels with warp-blocked strided chains + per-level input map (coalesced streams): parity, then
# the large sweeps and the chain-length knobs.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "affine or known_discrete or msm_matches or window_sizes or fold" > gpurun_out/pytest_r2o.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2o.log
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 5 >> gpurun_out/$out 2>> gpurun_out/r2o.err; }
: > gpurun_out/sweep_msm_bls12_381_g1_r2o.jsonl; : > gpurun_out/sweep_msm_bn254_g1_r2o.jsonl; : > gpurun_out/tune_pair_r2o.jsonl
sw sweep_msm_bls12_381_g1_r2o.jsonl msm --curve bls12_381 --min 21 --max 24
sw sweep_msm_bn254_g1_r2o.jsonl msm --curve bn254 --min 22 --max 24
for m in 32 128; do sw tune_pair_r2o.jsonl msm --curve bls12_381 --min 24 --max 24 --opt msm_pair_m=$m; done
for m2 in 8 16; do sw tune_pair_r2o.jsonl msm --curve bls12_381 --min 24 --max 24 --opt msm_pair_m2=$m2; done
python - <<'PY'
import json, glob
for f in ["sweep_msm_bls12_381_g1_r2o", "sweep_msm_bn254_g1_r2o", "tune_pair_r2o"]:
    for l in open("gpurun_out/%s.jsonl" % f):
        r = json.loads(l); print(f, r["log_n"], round(r["ms"], 3), r.get("window_bits"), r.get("opts"), {k: round(v, 2) for k, v in (r.get("stage_ms") or {}).items()}, r.get("check"))
PY
python tools/profile_target.py 24 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/traffic_r2o.csv python tools/profile_target.py 24 > gpurun_out/ncu_traffic_r2o.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/traffic_r2o.csv")) if len(r) > 10 and r[0].isdigit()]
byid = collections.OrderedDict()
for r in rows:
    d = byid.setdefault(r[0], {"name": r[4]}); d[r[12]] = float(r[14]); d["u_" + r[12]] = r[13]
ids = list(byid)
fin = [i for i in ids if "k_msm_final" in byid[i]["name"]]
lo, hi = ids.index(fin[-2]) + 1, ids.index(fin[-1]) + 1
agg = collections.OrderedDict()
for i in ids[lo:hi]:
    d = byid[i]; k = d["name"].split("(")[0].replace("void ", "").replace("zkm::", "")[:44]
    a = agg.setdefault(k, [0, 0, 0, 0]); a[0] += d.get("gpu__time_duration.sum", 0); a[1] += d.get("dram__bytes_read.sum", 0); a[2] += d.get("dram__bytes_write.sum", 0); a[3] += 1
print("units:", {k: v for k, v in byid[ids[lo]].items() if k.startswith("u_")})
for k, v in agg.items(): print("  %-46s t=%12.1f rd=%14.1f wr=%14.1f x%d" % (k, v[0], v[1], v[2], v[3]))
PY
tail -3 gpurun_out/r2o.err

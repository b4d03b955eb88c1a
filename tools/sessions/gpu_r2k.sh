#!/bin/bash
# GPU session r2k: r2i again after the lane fix (round-robin, idle lanes first), two-stage fold, aggregation only on skew-prone windows.
# the small / mid-size sweeps again.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_multidev.py -m gpu -q -x --maxfail=5 -k "msm or golden or kzg or concurrent or proving or multidev or abi or points" > gpurun_out/pytest_r2k.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r2k.log
: > gpurun_out/proxy_inflight_r2k.jsonl
for k in 1 2 3 4 6 8; do
  timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 96 --inflight $k >> gpurun_out/proxy_inflight_r2k.jsonl 2>> gpurun_out/r2k.err
done
python - <<'PY'
import json
for l in open("gpurun_out/proxy_inflight_r2k.jsonl"):
    r = json.loads(l); print(r["proofs_in_flight"], round(r["ms_per_proof"], 3), round(r["proofs_per_s"], 1), r["kernel_launches_per_proof"])
PY
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 5 > gpurun_out/$out 2>> gpurun_out/r2k.err; }
sw sweep_msm_bls12_381_g1_r2k.jsonl msm --curve bls12_381 --min 16 --max 24
sw sweep_msm_bls12_381_g1_witness_r2k.jsonl msm --curve bls12_381 --min 16 --max 22 --kind witness
sw sweep_msm_bls12_381_g1_pre_r2k.jsonl msm --curve bls12_381 --min 16 --max 18 --kind witness --precompute
sw sweep_msm_bw6_761_g1_r2k.jsonl msm --curve bw6_761 --min 16 --max 18
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_msm_*_r2k.jsonl")):
    for l in open(f):
        r = json.loads(l); print(f.split("/")[-1][10:-10], r["log_n"], round(r["ms"], 3), r.get("window_bits"), {k: round(v, 2) for k, v in (r.get("stage_ms") or {}).items()}, r.get("check"))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_proxy_r2k.csv python tools/groth16_proxy.py --log-n 16 --proofs 2 --inflight 1 --serial > gpurun_out/ncu_proxy_r2k.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_proxy_r2k.csv")) if len(r) > 10 and r[0].isdigit()]
names = [r[4] for r in rows]; times = [float(r[-1]) for r in rows]
qp = [i for i, n in enumerate(names) if "k_qap_pointwise" in n]
lo = qp[-1] - 12
agg = collections.OrderedDict(); cnt = collections.Counter()
for n, t in zip(names[lo:], times[lo:]):
    k = n.split("(")[0].replace("void ", "").replace("zkm::", "")[:60]
    agg[k] = agg.get(k, 0) + t; cnt[k] += 1
print("last proof: launches %d, sum of kernel time %.1f us" % (len(names) - lo, sum(times[lo:]) / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:22]:
    print("   %-62s %9.1f us x%d" % (k, v / 1e3, cnt[k]))
PY
tail -3 gpurun_out/r2k.err

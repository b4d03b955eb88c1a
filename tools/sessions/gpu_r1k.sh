#!/bin/bash
# GPU session r1k: level-0 x-coordinate array: parity, A/B at 2^22..2^24, then the default bench run.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_all_r1k.log 2>&1
echo "all rc=$?" | tee -a gpurun_out/pytest_all_r1k.log
tail -3 gpurun_out/pytest_all_r1k.log
: > gpurun_out/xarr_ab_r1k.jsonl
for x in 0 1; do
  timeout 300 python tools/sweep.py msm --curve bls12_381 --group 1 --min 22 --max 24 --reps 3 --opt msm_xarr=$x >> gpurun_out/xarr_ab_r1k.jsonl 2>> gpurun_out/sweep_r1k.err
done
python - <<'PY'
import json
for l in open("gpurun_out/xarr_ab_r1k.jsonl"):
    r = json.loads(l); print(r["log_n"], r.get("opts"), round(r["ms"], 2), r["stage_ms"])
PY
tail -3 gpurun_out/sweep_r1k.err
timeout 900 python bench.py > gpurun_out/bench_r1_k.json 2> gpurun_out/bench_r1_k.err
echo "bench rc=$?"; python -c "
import json
r = json.load(open('gpurun_out/bench_r1_k.json'))
print(r['value'], r['e2e']['value'], r['roofline']['frac'], r['msm_stage_ms'], r.get('ntt',{}).get('ms'), r.get('groth16_proxy',{}).get('proofs_per_s'))
"
tail -3 gpurun_out/bench_r1_k.err

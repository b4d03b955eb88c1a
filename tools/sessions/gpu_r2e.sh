#!/bin/bash
# GPU session r2e: per-tile four-step tables in the NTT (parity + timing), launch lists of mid-size MSMs (where does the tail go?)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_large.py -m gpu -q --maxfail=10 -k "ntt or golden or witness" > gpurun_out/pytest_r2e.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2e.log
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 5 > gpurun_out/$out 2>> gpurun_out/r2e.err; }
sw sweep_ntt_bls12_381_r2e.jsonl ntt --curve bls12_381 --min 16 --max 26
sw sweep_ntt_bw6_761_r2e.jsonl ntt --curve bw6_761 --min 20 --max 23
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*_r2e.jsonl")):
    for l in open(f):
        r = json.loads(l); print(f.split("/")[-1][6:-11], r["log_n"], round(r.get("ms", r.get("fft_ms")), 3), round(r.get("coset_ifft_ms", 0), 3), r.get("check"))
PY
for lg in 21 16; do
  python tools/profile_target.py $lg > gpurun_out/pt_$lg.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_2p${lg}_r2e.csv python tools/profile_target.py $lg > gpurun_out/ncu_pt_$lg.log 2>&1
done
python - <<'PY'
import csv, collections
for lg in (21, 16):
    rows = [r for r in csv.reader(open("gpurun_out/launches_2p%d_r2e.csv" % lg)) if len(r) > 10 and r[0].isdigit()]
    # the second MSM of the run: take the last launch of k_msm_final and walk back to the previous one
    names = [r[4] for r in rows]; times = [float(r[-1]) for r in rows]
    fin = [i for i, n in enumerate(names) if "k_msm_final" in n]
    lo, hi = fin[-2] + 1, fin[-1] + 1
    agg = collections.OrderedDict()
    for n, t in zip(names[lo:hi], times[lo:hi]):
        k = n.split("<")[0].split("(")[0].replace("void ", "")
        agg[k] = agg.get(k, 0) + t
    unit = rows[0][-2] if rows else ""
    print("2^%d MSM launches %d total %.3f" % (lg, hi - lo, sum(times[lo:hi])))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:14]:
        print("   %-40s %10.1f" % (k, v))
PY
tail -3 gpurun_out/r2e.err

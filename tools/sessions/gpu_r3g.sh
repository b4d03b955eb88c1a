#!/bin/bash
# GPU session r3g: last session of round 2 -- launch lists of one 2^21 and one 2^16 MSM on the final kernels (where the fixed
# tail of a shard / of a zkMember-sized MSM goes), then the whole GPU suite and smoke on the final library.
mkdir -p gpurun_out
for lg in 21 16; do
  python tools/profile_target.py $lg > gpurun_out/pt_$lg.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_2p${lg}_r3g.csv python tools/profile_target.py $lg > gpurun_out/ncu_pt_${lg}_r3g.log 2>&1
done
python - <<'PY'
import csv, collections
for lg in (21, 16):
    rows = [r for r in csv.reader(open("gpurun_out/launches_2p%d_r3g.csv" % lg)) if len(r) > 10 and r[0].isdigit()]
    names = [r[4] for r in rows]; times = [float(r[-1]) for r in rows]
    fin = [i for i, n in enumerate(names) if "k_msm_final" in n]
    lo, hi = fin[-2] + 1, fin[-1] + 1
    agg = collections.OrderedDict()
    for n, t in zip(names[lo:hi], times[lo:hi]):
        k = n.split("<")[0].split("(")[0].replace("void ", "").replace("zkm::", "")
        agg[k] = agg.get(k, 0) + t
    print("2^%d MSM: %d launches, %.1f us of kernel time" % (lg, hi - lo, sum(times[lo:hi]) / 1e3))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]:
        print("   %-40s %10.1f us" % (k, v / 1e3))
PY
timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/pytest_gpu_r3g.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r3g.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r3g.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r3g.log

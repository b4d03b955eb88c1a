#!/bin/bash
# GPU session r2g: where does a zkMember-sized Groth16 proof spend its time?  Throughput vs proofs in flight, the serial
# per-proof GPU time (ncu launch list), and host_wait modes.
mkdir -p gpurun_out
: > gpurun_out/proxy_inflight_r2g.jsonl
for k in 1 2 3 4 6 8 12; do
  timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 96 --inflight $k >> gpurun_out/proxy_inflight_r2g.jsonl 2>> gpurun_out/r2g.err
done
timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 48 --inflight 1 --serial >> gpurun_out/proxy_inflight_r2g.jsonl 2>> gpurun_out/r2g.err
timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 96 --inflight 4 --host-wait 2 >> gpurun_out/proxy_inflight_r2g.jsonl 2>> gpurun_out/r2g.err
timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 96 --inflight 4 --host-wait 1 >> gpurun_out/proxy_inflight_r2g.jsonl 2>> gpurun_out/r2g.err
python - <<'PY'
import json
for l in open("gpurun_out/proxy_inflight_r2g.jsonl"):
    r = json.loads(l); print(r["proofs_in_flight"], r["concurrent_msms"], round(r["ms_per_proof"], 3), round(r["proofs_per_s"], 1), r["kernel_launches_per_proof"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_proxy_r2g.csv python tools/groth16_proxy.py --log-n 16 --proofs 2 --inflight 1 --serial > gpurun_out/ncu_proxy_r2g.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_proxy_r2g.csv")) if len(r) > 10 and r[0].isdigit()]
names = [r[4] for r in rows]; times = [float(r[-1]) for r in rows]
# the last proof: from the last k_qap_pointwise back to the 6 k_ntt passes before it
qp = [i for i, n in enumerate(names) if "k_qap_pointwise" in n]
lo = qp[-1] - 12
agg = collections.OrderedDict(); cnt = collections.Counter()
for n, t in zip(names[lo:], times[lo:]):
    k = n.split("(")[0].replace("void ", "").replace("zkm::", "")[:60]
    agg[k] = agg.get(k, 0) + t; cnt[k] += 1
print("last proof: launches %d, sum of kernel time %.1f us" % (len(names) - lo, sum(times[lo:]) / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:25]:
    print("   %-62s %9.1f us x%d" % (k, v / 1e3, cnt[k]))
PY
tail -3 gpurun_out/r2g.err

#!/usr/bin/env python3
"""profiles/roofline_traffic.json from an ncu launch list with DRAM counters:

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
      --log-file gpurun_out/traffic_<tag>.csv python tools/profile_target.py 24
  python tools/ncu_traffic.py gpurun_out/traffic_<tag>.csv profiles/roofline_traffic.json <tag>

Takes the LAST 2^24 MSM and the LAST NTT of the run (tools/profile_target.py runs each twice: warm, measured) and sums
dram__bytes_read + dram__bytes_write per kernel: the `traffic` fields bench.py reports next to the algorithmic bytes."""
import collections
import csv
import json
import sys

src, dst, tag = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
rows = list(csv.reader(open(src)))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r)
col = {n: i for i, n in enumerate(rows[hdr])}
launches = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= col["Metric Value"]:
        continue
    lid = int(r[col["ID"]])
    d = launches.setdefault(lid, {"name": r[col["Kernel Name"]]})
    v = float(r[col["Metric Value"]].replace(",", ""))
    unit = r[col["Metric Unit"]]
    m = r[col["Metric Name"]]
    if m.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    if m.startswith("gpu__time"):
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)     # -> ms
    d[m] = v
ids = list(launches)
names = [launches[i]["name"] for i in ids]
fin = [k for k, n in enumerate(names) if "k_msm_final" in n]
lo, hi = fin[-2] + 1, fin[-1] + 1
short = lambda n: n.split("(")[0].replace("void ", "").replace("zkm::", "")
acc_kernels = ("k_pair_fwd", "k_pair_bwd", "k_inv_batch", "k_accum_affine", "k_build_xarr", "k_pair_map", "k_pair_lens")
per = collections.OrderedDict()
msm_total = msm_ms = acc_total = acc_ms = 0.0
for k in range(lo, hi):
    d = launches[ids[k]]
    b = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    t = d.get("gpu__time_duration.sum", 0.0)
    key = short(d["name"]).split("<")[0]
    e = per.setdefault(key, {"launches": 0, "gb": 0.0, "ms": 0.0})
    e["launches"] += 1
    e["gb"] += b / 1e9
    e["ms"] += t
    msm_total += b
    msm_ms += t
    if key in acc_kernels:
        acc_total += b
        acc_ms += t
ntt = [k for k, n in enumerate(names) if "k_ntt_pass" in n]
passes = 2
ntt_bytes = sum(launches[ids[k]].get("dram__bytes_read.sum", 0) + launches[ids[k]].get("dram__bytes_write.sum", 0) for k in ntt[-passes:])
ntt_ms = sum(launches[ids[k]].get("gpu__time_duration.sum", 0) for k in ntt[-passes:])
out = {
    "workload": "bls12_381_g1_msm_2p24",
    "source": "ncu launch list with DRAM counters of tools/profile_target.py 24 (%s): last MSM and last NTT of the run; profiles/%s"
              % (tag, src.split("/")[-1]),
    "bucket_accumulation_dram_bytes": acc_total,
    "bucket_accumulation_ms_under_ncu": acc_ms,
    "msm_dram_bytes": msm_total,
    "msm_ms_under_ncu": msm_ms,
    "msm_per_kernel": {k: {"launches": v["launches"], "gb": round(v["gb"], 3), "ms": round(v["ms"], 3)} for k, v in per.items()},
    "ntt_2p24_dram_bytes": ntt_bytes,
    "ntt_2p24_ms_under_ncu": ntt_ms,
}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))

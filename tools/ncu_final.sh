#!/bin/bash
# Round-end evidence: (1) ncu launch list of the bench command itself, (2) --set full captures of the
# dominant kernels on tools/profile_target.py, exported to CSV on the GPU box (the .ncu-rep files with
# imported source exceed the 64 MiB return limit).  Usage: tools/ncu_final.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-proxy --skip-precompute > $OUT/bench_plain_$TAG.json 2> $OUT/bench_plain_$TAG.err || { echo "plain bench failed"; tail -5 $OUT/bench_plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-proxy --skip-precompute > $OUT/ncu_bench_$TAG.log 2>&1
python tools/profile_target.py 24 > $OUT/pt_plain_$TAG.log 2>&1 || { echo "plain target failed"; exit 1; }
# DRAM traffic of every launch of the fixed workload -> profiles/roofline_traffic.json (bench.py's `traffic` fields)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file $OUT/traffic_$TAG.csv python tools/profile_target.py 24 > $OUT/ncu_traffic_$TAG.log 2>&1
python tools/ncu_traffic.py $OUT/traffic_$TAG.csv $OUT/roofline_traffic_$TAG.json $TAG > $OUT/ncu_traffic_summary_$TAG.txt 2>&1; tail -5 $OUT/ncu_traffic_summary_$TAG.txt
cap() { # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/prof_$1 -f \
      python tools/profile_target.py 24 > $OUT/ncu_$1_$TAG.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > $OUT/ncu_raw_$1_$TAG.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > /tmp/src_$1.csv 2>/dev/null
  gzip -c /tmp/src_$1.csv > $OUT/ncu_source_$1_$TAG.csv.gz
  ls -la $OUT/ncu_raw_$1_$TAG.csv $OUT/ncu_source_$1_$TAG.csv.gz
}
cap pairbwd "k_pair_bwd" 4 1
cap pairfwd "k_pair_fwd" 4 1
cap accum "k_accum_affine" 1 1
cap reduce "k_bucket_reduce" 1 1
cap ntt "k_ntt_pass" 2 2

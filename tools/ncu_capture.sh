#!/bin/bash
# ncu --set full captures of the dominant kernels on the fixed profiling workload
# (tools/profile_target.py), exported to CSV on the GPU box so that only small text files
# travel back (the .ncu-rep with imported source is > 64 MiB).  Usage: tools/ncu_capture.sh <tag> [log_n]
set -u
TAG=${1:-r1}
LOGN=${2:-24}
OUT=gpurun_out
mkdir -p $OUT
python tools/profile_target.py $LOGN > $OUT/pt_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/pt_plain_$TAG.log; exit 1; }
cap() { # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/prof_$1 -f \
      python tools/profile_target.py $LOGN > $OUT/ncu_$1_$TAG.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > $OUT/ncu_raw_$1_$TAG.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > /tmp/src_$1.csv 2>/dev/null
  # keep the source page small: top 400 lines by sampled stalls are extracted by the reader; gzip the full page
  gzip -c /tmp/src_$1.csv > $OUT/ncu_source_$1_$TAG.csv.gz
  ls -la /tmp/prof_$1.ncu-rep $OUT/ncu_raw_$1_$TAG.csv $OUT/ncu_source_$1_$TAG.csv.gz
}
cap accum "k_accum_affine" 1 1
cap ntt "k_ntt_pass" 3 3
cap reduce "k_bucket_reduce" 1 1
cap scatter "k_msm_digits" 14 2

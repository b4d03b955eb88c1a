#!/bin/bash
# ncu --set full on the batched-affine kernels (second level: contiguous sources). Usage: tools/ncu_capture2.sh <tag>
set -u
TAG=${1:-r1b}
OUT=gpurun_out
mkdir -p $OUT
python tools/profile_target.py 24 > $OUT/pt_plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/pt_plain_$TAG.log; exit 1; }
cap() {
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/prof_$1 -f \
      python tools/profile_target.py 24 > $OUT/ncu_$1_$TAG.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > $OUT/ncu_raw_$1_$TAG.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > /tmp/src_$1.csv 2>/dev/null
  gzip -c /tmp/src_$1.csv > $OUT/ncu_source_$1_$TAG.csv.gz
  ls -la $OUT/ncu_raw_$1_$TAG.csv $OUT/ncu_source_$1_$TAG.csv.gz
}
cap pairbwd "k_pair_bwd" 4 2
cap pairfwd "k_pair_fwd" 4 2

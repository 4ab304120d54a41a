#!/bin/bash
OUT=gpurun_out; TAG=r01_v41
cap() {
  timeout 600 ncu --set full --clock-control none --profile-from-start off -k "regex:$2" -s $3 -c $4 -f -o $OUT/${TAG}_$1 python tools/profile_step.py > $OUT/${TAG}_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i $OUT/${TAG}_$1.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_$1_raw.csv.gz
  rm -f $OUT/${TAG}_$1.ncu-rep
}
cap wgraddb 'wgrad64_db_kernel' 1 2
cap lin     'lin_tc_kernel'     10 8
ls -la $OUT/${TAG}_*

#!/bin/bash
# One ncu --set full capture per kernel class of the bench-configuration train step (run on the GPU box, after a plain run
# of the same program has exited 0).  Usage: tools/capture_profiles.sh <tag>   -> gpurun_out/<tag>_<class>.ncu-rep
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python tools/profile_step.py > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
cap() {  # name regex skip count [extra]
  timeout 600 ncu --set full --clock-control none --profile-from-start off -k "regex:$2" -s $3 -c $4 ${5:-} -f -o $OUT/${TAG}_$1 \
    python tools/profile_step.py > $OUT/${TAG}_$1.log 2>&1
  echo "$1 rc=$?"
  # only the exported tables travel back (gpurun merges at most 64 MiB); the conv report itself is kept for the source page
  ncu -i $OUT/${TAG}_$1.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_$1_raw.csv.gz
  if [ "$1" != "conv" ]; then rm -f $OUT/${TAG}_$1.ncu-rep; fi
}
cap conv    'conv64_tc_kernel'      2 2 "--import-source on"
cap wgrad   'wgrad64_tc_kernel'     1 1
cap lin     'lin_tc_kernel'         10 8
cap linwg   'lin_wgrad_t(c|ma)_kernel' 4 5
cap attn    'attn_(fwd|bwd)_tc'     1 1
cap attnb   'attn_bwd_tc'           1 1
cap ew      'ln_ct|ln64|ct_reduce|add_kernel|adamw' 4 10
cap lstm    'lstm128'               0 2
cap head    'logits_tc'             0 3
ls -la $OUT/${TAG}_*

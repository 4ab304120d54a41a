#!/bin/bash
# Round-2 final ncu evidence (run under gpurun on one B200, after the plain program has exited 0):
#   1. launch list of one bench-configuration step (per-launch gpu__time_duration)
#   2. --set full captures of the kernels that changed since v47 (LSTM cluster recurrences, fp16 attention, tensor-map fed weight
#      gradient) + the roofline kernel (for roofline.traffic)
OUT=gpurun_out; TAG=${1:-r02_v55}
python tools/profile_step.py > $OUT/${TAG}_step_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches.csv \
    python tools/profile_step.py > $OUT/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
cap() {   # name, kernel regex, skip, count
  local name=$1 rx=$2 skip=$3 cnt=$4
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$rx" -s $skip -c $cnt -f -o $OUT/${TAG}_$name \
      python tools/profile_step.py > $OUT/${TAG}_$name.log 2>&1
  echo "$name rc=$?"
  ncu -i $OUT/${TAG}_$name.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_${name}_raw.csv.gz
  rm -f $OUT/${TAG}_$name.ncu-rep
}
cap conv   'conv64_tc_kernel'          2 2
cap lstm   'lstm128_(fwd|bwd)_c2'      0 2
cap attnf  'attn_fwd_h_kernel'         2 1
cap attnb  'attn_bwd_h_kernel'         1 1
cap linwg  'lin_wgrad_tma_kernel'      0 6
cap lin    'lin_tc_kernel<3, 0, (13|73|1),' 6 3
ls -la $OUT/${TAG}_*

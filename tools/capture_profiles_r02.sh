#!/bin/bash
# Round-2 ncu evidence (run under gpurun on one B200, after the plain programs have exited 0):
#   1. launch list of one bench-configuration step (per-launch gpu__time_duration)
#   2. --set full captures of the kernels that changed this round + the roofline kernel (for roofline.traffic)
OUT=gpurun_out; TAG=${1:-r02_v47}
python tools/profile_step.py > $OUT/${TAG}_step_plain.log 2>&1 || exit 1
python tools/time_head.py 4096 2560 2 > $OUT/${TAG}_head_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches.csv \
    python tools/profile_step.py > $OUT/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_head_launches.csv \
    python tools/time_head.py 4096 2560 1 > $OUT/${TAG}_head_launches.log 2>&1
echo "head launch list rc=$?"
cap() {   # name, kernel regex, skip, count, program...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on "${PFS[@]}" -k "regex:$rx" -s $skip -c $cnt -f -o $OUT/${TAG}_$name "$@" > $OUT/${TAG}_$name.log 2>&1
  echo "$name rc=$?"
  ncu -i $OUT/${TAG}_$name.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_${name}_raw.csv.gz
  rm -f $OUT/${TAG}_$name.ncu-rep
}
PFS=(--profile-from-start off)
cap conv   'conv64_tc_kernel'        2 2 python tools/profile_step.py
cap lnb    'lin_tc_kernel<3, 0, 128' 1 2 python tools/profile_step.py
cap lin    'lin_tc_kernel<3, 0, (13|73|1),' 6 4 python tools/profile_step.py
cap attn   'attn_(fwd|bwd)_tc_kernel' 2 2 python tools/profile_step.py
PFS=()
cap head   'logits_tc_kernel|epack'  12 10 python tools/time_head.py 4096 2560 1
ls -la $OUT/${TAG}_*

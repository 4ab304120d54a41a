#!/bin/bash
# ncu --set full capture of the 64-column-tile similarity kernel inside one bench-configuration step (B = 256: 8 CTAs per launch)
OUT=gpurun_out; TAG=${1:-r02_v62}
python tools/profile_step.py > $OUT/${TAG}_step_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:logits_tc_kernel" -c 7 -f -o $OUT/${TAG}_head \
    python tools/profile_step.py > $OUT/${TAG}_head.log 2>&1
echo "head rc=$?"
ncu -i $OUT/${TAG}_head.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_head_raw.csv.gz
rm -f $OUT/${TAG}_head.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches.csv \
    python tools/profile_step.py > $OUT/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
cat $OUT/${TAG}_step_plain.log

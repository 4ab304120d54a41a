#!/bin/bash
# Round-2 closing evidence at build v58 (run under gpurun on one B200): the plain bench line first, then -- only after it has
# exited 0 -- the ncu launch list of one bench-configuration step and --set full captures of the roofline kernel and the
# attention backward (the kernel changed since the v55 captures).
OUT=gpurun_out; TAG=${1:-r02_v58}
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || { tail -5 $OUT/${TAG}_bench.err; exit 1; }
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
cap attnb  'attn_bwd_h_kernel'         1 1
ls -la $OUT/${TAG}_*

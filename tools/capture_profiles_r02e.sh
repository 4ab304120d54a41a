#!/bin/bash
# Round-2 closing evidence at the final build (run under gpurun on one B200): the plain bench line first, then -- only after it has
# exited 0 -- the ncu launch list of one bench-configuration step.
OUT=gpurun_out; TAG=${1:-r02_v63}
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || { tail -5 $OUT/${TAG}_bench.err; exit 1; }
python tools/profile_step.py > $OUT/${TAG}_step_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/${TAG}_launches.csv \
    python tools/profile_step.py > $OUT/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
cat $OUT/${TAG}_step_plain.log

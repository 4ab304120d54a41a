#!/bin/bash
# ncu --set full capture of the attention kernels inside one bench-configuration step (run under gpurun after a plain run exited 0)
OUT=gpurun_out; TAG=${1:-r02_v54}
python tools/profile_step.py > $OUT/${TAG}_step_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:attn_(fwd|bwd)_h_kernel" -s 2 -c 2 -f -o $OUT/${TAG}_attn python tools/profile_step.py > $OUT/${TAG}_attn.log 2>&1
echo "attn rc=$?"
ncu -i $OUT/${TAG}_attn.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_attn_raw.csv.gz
ncu -i $OUT/${TAG}_attn.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > $OUT/${TAG}_attn_source.csv.gz
rm -f $OUT/${TAG}_attn.ncu-rep
ls -la $OUT/${TAG}_*

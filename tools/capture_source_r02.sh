#!/bin/bash
# source-level stall attribution of one token-GEMM launch (FFN2 forward: K = 256 -> N = 64, bias + dropout + residual epilogue)
OUT=gpurun_out; TAG=r02_v51
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
  -k "regex:lin_tc_kernel<.int.3, .int.0, .int.73, .int.4" -s 2 -c 1 -f -o $OUT/${TAG}_ffn2 python tools/profile_step.py > $OUT/${TAG}_ffn2.log 2>&1
echo "rc=$?"
ncu -i $OUT/${TAG}_ffn2.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > $OUT/${TAG}_ffn2_sass.csv.gz
ncu -i $OUT/${TAG}_ffn2.ncu-rep --page source --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_ffn2_src.csv.gz
ncu -i $OUT/${TAG}_ffn2.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_ffn2_raw.csv.gz
rm -f $OUT/${TAG}_ffn2.ncu-rep
ls -la $OUT/${TAG}_ffn2*

OUT=gpurun_out; TAG=r02_v47
cap() {
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k "regex:$rx" -s $skip -c $cnt -f -o $OUT/${TAG}_$name "$@" > $OUT/${TAG}_$name.log 2>&1
  echo "$name rc=$?"
  ncu -i $OUT/${TAG}_$name.ncu-rep --page raw --csv 2>/dev/null | gzip -9 > $OUT/${TAG}_${name}_raw.csv.gz
  rm -f $OUT/${TAG}_$name.ncu-rep
}
cap lnb 'lin_tc_kernel<.int.3, .int.0, .int.128' 1 2 python tools/profile_step.py
cap lin 'lin_tc_kernel<.int.3, .int.0, .int.(13|73|1).,' 6 4 python tools/profile_step.py
ls -la $OUT/${TAG}_l*

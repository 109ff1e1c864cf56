#!/bin/bash
# One gpurun call: plain run, ncu launch list of the same command, then --set full captures of the hot kernels.
# usage: tools/ncu_capture.sh <tag> [workload]
TAG=${1:-r1}
WL=${2:-C2}
CMD="python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
[ -n "$NO_LIST" ] || ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_${TAG}_$1 $CMD > gpurun_out/ncu_$1_$TAG.log 2>&1
  # gpurun_out/ is capped at 64 MiB: keep the CSV pages, drop the report (except the dominant kernel's)
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/raw_${TAG}_$1.csv 2>/dev/null
  ncu -i gpurun_out/prof_${TAG}_$1.ncu-rep --page source --csv > gpurun_out/source_${TAG}_$1.csv 2>/dev/null
  [ "$1" = "zw" ] || rm -f gpurun_out/prof_${TAG}_$1.ncu-rep
}
WHICH=${3:-"zw f2 i2 f3 i1"}
for k in $WHICH; do
  case $k in
    zw) cap zw 'pass_z(w|16)' 60 ;;
    f2) cap f2 pass_f2_kernel 80 ;;
    i2) cap i2 pass_i2_kernel 40 ;;
    f3) cap f3 pass_f3_kernel 130 ;;
    i1) cap i1 pass_i1_kernel 40 ;;
  esac
done
ls -la gpurun_out

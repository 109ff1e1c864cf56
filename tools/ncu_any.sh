#!/bin/bash
# ncu --set full captures of the any-length passes on the BOSS production mesh (one gpurun call).
# usage: tools/ncu_any.sh <tag> [kernels...]   kernels from: f2 f3 i2 i1 z zr
TAG=${1:-r2q}; shift
WHICH=${@:-"f2 f3 z"}
CMD="python bench.py --workload BOSSPK --fft 1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --inflight 1"
mkdir -p gpurun_out/$TAG
$CMD > gpurun_out/$TAG/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/$TAG/plain.log; exit 1; }
cap() {  # name regex skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/$TAG/prof_$1 $CMD > gpurun_out/$TAG/ncu_$1.log 2>&1
  ncu -i gpurun_out/$TAG/prof_$1.ncu-rep --page raw --csv > gpurun_out/$TAG/raw_$1.csv 2>/dev/null
  ncu -i gpurun_out/$TAG/prof_$1.ncu-rep --page source --csv > gpurun_out/$TAG/source_$1.csv 2>/dev/null
  rm -f gpurun_out/$TAG/prof_$1.ncu-rep
}
for k in $WHICH; do
  case $k in
    f2) cap f2 gpass_f2_kernel 40 ;;
    f3) cap f3 gpass_f3_kernel 80 ;;
    i2) cap i2 gpass_i2_kernel 10 ;;
    i1) cap i1 gpass_i1_kernel 10 ;;
    z) cap z 'gpass_z_kernel' 5 ;;
  esac
done
ls -la gpurun_out/$TAG

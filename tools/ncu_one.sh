#!/bin/bash
# One ncu --set full capture of one kernel of a bench.py workload: tools/ncu_one.sh <tag> <name> <kernel regex> <skip> [bench args...]
TAG=$1; NAME=$2; RE=$3; SKIP=$4; shift 4
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --inflight 1 $@"
mkdir -p gpurun_out/$TAG
$CMD > gpurun_out/$TAG/plain_$NAME.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/$TAG/plain_$NAME.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c 1 -f -o gpurun_out/$TAG/prof_$NAME $CMD > gpurun_out/$TAG/ncu_$NAME.log 2>&1
ncu -i gpurun_out/$TAG/prof_$NAME.ncu-rep --page raw --csv > gpurun_out/$TAG/raw_$NAME.csv 2>/dev/null
ncu -i gpurun_out/$TAG/prof_$NAME.ncu-rep --page source --csv > gpurun_out/$TAG/source_$NAME.csv 2>/dev/null
rm -f gpurun_out/$TAG/prof_$NAME.ncu-rep
ls -la gpurun_out/$TAG

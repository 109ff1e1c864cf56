#!/bin/bash
# ncu --set full captures of the bispectrum Fisher (C4) hot kernels: batched inverse z pass and the DMMA Gram
CMD="python bench.py --workload C4 --steps 1 --warmup 1"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_c4.log 2>&1 || { echo plain failed; tail -5 gpurun_out/plain_c4.log; exit 1; }
cap() {
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_c4_$1 $CMD > gpurun_out/ncu_c4_$1.log 2>&1
  ncu -i gpurun_out/prof_c4_$1.ncu-rep --page raw --csv > gpurun_out/raw_c4_$1.csv 2>/dev/null
  ncu -i gpurun_out/prof_c4_$1.ncu-rep --page source --csv > gpurun_out/source_c4_$1.csv 2>/dev/null
  rm -f gpurun_out/prof_c4_$1.ncu-rep
}
cap zinv 'pass_zw_kernel' ${ZSKIP:-1500}
cap gram 'gram_dmma_kernel' 3
ls -la gpurun_out | head -30

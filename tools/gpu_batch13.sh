#!/bin/bash
# round 2, batch 13 (1 GPU): which instruction classes share issue resources (tools/micro/pipes.cu), the sorting network
# alone (tools/micro/sortbench.cu), and the register kernel with phases switched off (timing experiments, SSC_DIAG)
set -u
O=gpurun_out/b13
mkdir -p $O
./tools/micro/pipes > $O/pipes.txt 2>&1; cat $O/pipes.txt
timeout 120 ./tools/micro/sortbench > $O/sortbench.txt 2>&1; head -6 $O/sortbench.txt
for v in default diag4 diag5 diag6 diag7; do
  lib=$PWD/build/variants/libssc_$v.so; [ $v = default ] && lib=$PWD/siril-0.9_b200/libsiril_stack_cuda.so
  SSC_LIB=$lib timeout 200 python tools/profile_step.py --config cfg3 --steps 4 > $O/step_$v.txt 2>&1
  echo "$v: $(tail -1 $O/step_$v.txt)"
done

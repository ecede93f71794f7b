#!/bin/bash
# round 2, batch 14 (1 GPU): staggered start of the register kernel's resident blocks / warps (SSC_STAGGER_NS, SSC_STAGGER_WARP_NS)
set -u
O=gpurun_out/b14
mkdir -p $O
for cfg in "0 0" "2300 0" "4600 0" "1200 0" "0 600" "2300 600" "7000 0" "3000 750"; do
  set -- $cfg
  SSC_STAGGER_NS=$1 SSC_STAGGER_WARP_NS=$2 timeout 200 python tools/profile_step.py --config cfg3 --steps 4 > $O/step_$1_$2.txt 2>&1
  echo "stagger $1 warp $2: $(tail -1 $O/step_$1_$2.txt)"
done

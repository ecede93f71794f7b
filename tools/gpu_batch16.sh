#!/bin/bash
# round 2, batch 16 (1 GPU): cfg4 slice with 4 and with 3 resident blocks per SM (SSC_MID_SMEM_PAD), cheap normalization on / off
set -u
for pad in 0 2048 8192; do for ch in 1 0; do
echo "pad $pad cheap $ch: $(SSC_MID_SMEM_PAD=$pad SSC_CHEAP_NORM=$ch python tools/profile_step.py --config cfg4 --rows 440 --steps 3 2>&1 | tail -1)"
done; done

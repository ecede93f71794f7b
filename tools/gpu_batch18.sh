#!/bin/bash
# round 2, batch 18 (1 GPU): full ncu capture of the mid kernel on a cfg4 slice (440 rows = 1/8 of the image)
set -u
O=gpurun_out/b18
mkdir -p $O
python tools/profile_step.py --config cfg4 --rows 440 --steps 2 > $O/plain.log 2>&1 || { cat $O/plain.log; exit 1; }
tail -1 $O/plain.log
ncu --set full --clock-control none --import-source on -k regex:k_stack_mid -s 1 -c 1 -o $O/prof_cfg4_mid python tools/profile_step.py --config cfg4 --rows 440 --steps 2 > $O/ncu.log 2>&1
tail -3 $O/ncu.log; ls -la $O

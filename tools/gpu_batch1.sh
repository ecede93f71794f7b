#!/bin/bash
# round 2, batch 1: all -m gpu tests, every BASELINE config at N=1, launch lists + full ncu captures of
# the kernels that had no profile yet
set -u
O=gpurun_out/b1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do
  st=10; [ $c = cfg4 ] && st=5; [ $c = cfg5 ] && st=3
  timeout 900 python bench.py --config $c --steps $st --warmup 3 > $O/bench_${c}_n1.json 2> $O/bench_${c}_n1.err
  echo "$c exit $?"; cut -c1-400 $O/bench_${c}_n1.json
done
timeout 600 python bench.py --impl reference --config cfg3 --steps 2 --warmup 1 > $O/bench_cfg3_reference.json 2> $O/bench_cfg3_reference.err
# launch lists (cold-cache, serialised): shares per step
for c in cfg1 cfg2 cfg4; do
  python tools/profile_step.py --config $c --steps 2 > $O/plain_$c.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$c.csv python tools/profile_step.py --config $c --steps 2 > $O/ncu_l_$c.log 2>&1
done
python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/plain_cfg5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_cfg5.csv python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/ncu_l_cfg5.log 2>&1
# full captures of the dominant kernels (1 launch each, slices to keep replays short)
ncu --set full --clock-control none --import-source on -k regex:k_stack_mid -s 1 -c 1 -o $O/prof_cfg4_mid python tools/profile_step.py --config cfg4 --rows 440 --steps 2 > $O/ncu_f_cfg4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_stack_mid|k_tail_spilled' -s 2 -c 2 -o $O/prof_cfg5_fit python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/ncu_f_cfg5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_stack_fast -s 1 -c 1 -o $O/prof_cfg2_median python tools/profile_step.py --config cfg2 --steps 2 > $O/ncu_f_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_sum' -s 2 -c 2 -o $O/prof_cfg1_sum python tools/profile_step.py --config cfg1 --steps 2 > $O/ncu_f_cfg1.log 2>&1
ls -la $O

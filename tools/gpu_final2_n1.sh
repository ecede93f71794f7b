#!/bin/bash
# round 2, second set of final single-GPU measurements (after the mid / warp / K5 kernel changes): all -m gpu tests, every BASELINE
# config (bench line with roofline, e2e, cpu_baseline), launch lists of cfg3 / cfg4 / cfg5, full ncu capture of the warp kernel
set -u
O=gpurun_out/final2_n1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
for c in cfg3 cfg1 cfg2 cfg4 cfg5; do
  st=10; [ $c = cfg4 ] && st=5; [ $c = cfg5 ] && st=5
  timeout 600 python bench.py --config $c --steps $st --warmup 3 > $O/bench_${c}_n1.json 2> $O/bench_${c}_n1.err
  echo "$c exit $?"; cut -c1-170 $O/bench_${c}_n1.json
done
for c in cfg3 cfg4; do
  python tools/profile_step.py --config $c --steps 2 > $O/plain_$c.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$c.csv python tools/profile_step.py --config $c --steps 2 > $O/ncu_l_$c.log 2>&1
done
python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/plain_cfg5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_cfg5.csv python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/ncu_l_cfg5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_stack_warp' -s 1 -c 1 -o $O/prof_cfg5_warp python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/ncu_f_cfg5.log 2>&1
cat $O/plain_cfg5.log | tail -1
ls $O | wc -l

#!/bin/bash
# round 2, final single-GPU measurements: all -m gpu tests, every BASELINE config (bench line with roofline, e2e,
# cpu_baseline), the reference arm, launch lists of every config, full ncu capture of the warp kernel
set -u
O=gpurun_out/final_n1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
for c in cfg1 cfg2 cfg3 cfg4 cfg5; do
  st=10; [ $c = cfg4 ] && st=5; [ $c = cfg5 ] && st=5
  timeout 900 python bench.py --config $c --steps $st --warmup 3 > $O/bench_${c}_n1.json 2> $O/bench_${c}_n1.err
  echo "$c exit $?"; cut -c1-160 $O/bench_${c}_n1.json
done
timeout 900 python bench.py --impl reference --config cfg3 --steps 2 --warmup 1 > $O/bench_cfg3_reference.json 2> $O/bench_cfg3_reference.err
cut -c1-200 $O/bench_cfg3_reference.json
for c in cfg1 cfg2 cfg3 cfg4; do
  python tools/profile_step.py --config $c --steps 2 > $O/plain_$c.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$c.csv python tools/profile_step.py --config $c --steps 2 > $O/ncu_l_$c.log 2>&1
done
python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/plain_cfg5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_cfg5.csv python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/ncu_l_cfg5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_stack_warp|k_fit_handback' -s 2 -c 2 -o $O/prof_cfg5_warp python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/ncu_f_cfg5.log 2>&1
cat $O/plain_cfg5.log
ls $O | head -40

#!/bin/bash
# round 2, batch 3: all -m gpu tests; cfg3 (pruned network, packed-pair histogram) and cfg5 (branch-free fit) benches
set -u
O=gpurun_out/b3
mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
for c in cfg3 cfg5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu > $O/bench_${c}_n1.json 2> $O/bench_${c}_n1.err
  echo "$c exit $?"; cut -c1-250 $O/bench_${c}_n1.json
done
for n in 32 64 200; do python tools/profile_step.py --config cfg3 --rows 500 --rejection 5 --frames $n --steps 2; done > $O/fit_other.log 2>&1
cat $O/fit_other.log
python tools/profile_step.py --config cfg3 --rows 1000 --steps 3 > $O/plain_cfg3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_histogram|k_stack_fast' -s 3 -c 2 -o $O/prof_cfg3 python tools/profile_step.py --config cfg3 --rows 1000 --steps 3 > $O/ncu_f_cfg3.log 2>&1
cat $O/plain_cfg3.log

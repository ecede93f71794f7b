#!/bin/bash
# round 2, batch 2: all -m gpu tests with the warp kernel (linear fit), cfg5 bench, ncu of k_stack_warp<32>
set -u
O=gpurun_out/b2
mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 python bench.py --config cfg5 --steps 5 --warmup 3 > $O/bench_cfg5_n1.json 2> $O/bench_cfg5_n1.err
echo "cfg5 exit $?"; cut -c1-300 $O/bench_cfg5_n1.json
python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/plain_cfg5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_stack_warp' -s 1 -c 1 -o $O/prof_cfg5_warp python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/ncu_f_cfg5.log 2>&1
cat $O/plain_cfg5.log
for n in 64 200; do python tools/profile_step.py --config cfg3 --rows 500 --rejection 5 --frames $n --steps 2; done > $O/fit_other.log 2>&1
cat $O/fit_other.log

#!/bin/bash
# round 2, batch 5 (1 GPU): all -m gpu tests (file formats, preprocess, savefits); cfg3 end to end with the landing
# under the upload; ncu of the branch-free warp kernel
set -u
O=gpurun_out/b5
mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
SSC_TRACE=1 timeout 600 python bench.py --config cfg3 --steps 5 --warmup 3 --no-cpu > $O/bench_cfg3_n1.json 2> $O/bench_cfg3_n1.err
echo "cfg3 n1 exit $?"; cut -c1-200 $O/bench_cfg3_n1.json; grep "ssc trace" $O/bench_cfg3_n1.err | tail -12
python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/plain_cfg5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_stack_warp' -s 1 -c 1 -o $O/prof_cfg5_warp python tools/profile_step.py --config cfg5 --rows 152 --steps 2 > $O/ncu_f_cfg5.log 2>&1
cat $O/plain_cfg5.log

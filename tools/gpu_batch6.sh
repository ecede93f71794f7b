#!/bin/bash
# round 2, batch 6 (1 GPU): all -m gpu tests; cfg3 end to end (plan trace); cfg5 with the warp-autonomous kernel
set -u
O=gpurun_out/b6
mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
SSC_TRACE=1 timeout 600 python bench.py --config cfg3 --steps 5 --warmup 3 --no-cpu > $O/bench_cfg3_n1.json 2> $O/bench_cfg3_n1.err
echo "cfg3 n1 exit $?"; cut -c1-200 $O/bench_cfg3_n1.json; grep "ssc trace" $O/bench_cfg3_n1.err | tail -22
timeout 600 python bench.py --config cfg5 --steps 5 --warmup 3 --no-cpu > $O/bench_cfg5_n1.json 2> $O/bench_cfg5_n1.err
echo "cfg5 exit $?"; cut -c1-200 $O/bench_cfg5_n1.json
for n in 32 64 200; do python tools/profile_step.py --config cfg3 --rows 500 --rejection 5 --frames $n --steps 2; done > $O/fit_other.log 2>&1
cat $O/fit_other.log

#!/bin/bash
# round 2, batch 7 (1 GPU): all -m gpu tests after the K3 dynamic tile scheduling / plan reuse; cfg3 bench
set -u
O=gpurun_out/b7
mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
SSC_TRACE=1 timeout 600 python bench.py --config cfg3 --steps 10 --warmup 3 > $O/bench_cfg3_n1.json 2> $O/bench_cfg3_n1.err
echo "cfg3 n1 exit $?"; cut -c1-200 $O/bench_cfg3_n1.json; grep "ssc trace" $O/bench_cfg3_n1.err | tail -6
python tools/profile_step.py --config cfg3 --rows 500 --steps 4 > $O/plain_cfg3_500.log 2>&1; cat $O/plain_cfg3_500.log

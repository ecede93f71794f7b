#!/bin/bash
# round 2, batch 11 (1 GPU): all -m gpu tests after the deferred hand-back / pull exchange; cfg5 and cfg3 end to end
set -u
O=gpurun_out/b11
mkdir -p $O
timeout 1700 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
for c in cfg5 cfg3; do
SSC_TRACE=1 timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu --e2e-steps 3 > $O/bench_${c}_n1.json 2> $O/bench_${c}_n1.err
echo "$c n1 exit $?"; cut -c1-200 $O/bench_${c}_n1.json; grep "ssc trace" $O/bench_${c}_n1.err | grep -v unit | tail -6
done

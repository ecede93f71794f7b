#!/bin/bash
# round 2, batch 10 (1 GPU): cfg5 / cfg4 operator traces with per-unit kernel timing
set -u
O=gpurun_out/b10
mkdir -p $O
for c in cfg5 cfg4; do
SSC_TRACE=1 timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu --e2e-steps 2 > $O/bench_${c}_n1.json 2> $O/bench_${c}_n1.err
echo "$c n1 exit $?"; cut -c1-200 $O/bench_${c}_n1.json; grep "ssc trace" $O/bench_${c}_n1.err | tail -16
done

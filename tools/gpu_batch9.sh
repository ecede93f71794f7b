#!/bin/bash
# round 2, batch 9 (2 GPUs): pull-based exchange; traces of the multi-device operator; cfg4 / cfg5 at N=1
set -u
O=gpurun_out/b9
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_operator_paths.py tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
run() { # config gpus steps
  SSC_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29541 bench.py --config $1 --gpus $2 --steps $3 --warmup 3 --no-cpu > $O/bench_$1_n$2.json 2> $O/bench_$1_n$2.err
  echo "$1 n$2 exit $?"; cut -c1-200 $O/bench_$1_n$2.json
}
run cfg3 2 10
run cfg5 2 3
run cfg1 2 10
for c in cfg4 cfg5; do
SSC_TRACE=1 timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu > $O/bench_${c}_n1.json 2> $O/bench_${c}_n1.err
echo "$c n1 exit $?"; cut -c1-200 $O/bench_${c}_n1.json
done

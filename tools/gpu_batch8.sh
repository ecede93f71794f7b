#!/bin/bash
# round 2, batch 8 (8 GPUs): the multi-device operator tests on 8 devices, every BASELINE config at N=8,
# cfg3 with both exchanges
set -u
O=gpurun_out/b8
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_operator_paths.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
run() { # config exchange steps
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --config $1 --gpus 8 --steps $3 --warmup 3 --no-cpu --exchange $2 > $O/bench_$1_n8_$2.json 2> $O/bench_$1_n8_$2.err
  echo "$1 n8 $2 exit $?"; cut -c1-220 $O/bench_$1_n8_$2.json
}
run cfg3 peer 10
run cfg3 nccl 10
run cfg5 peer 5
run cfg4 peer 5
run cfg2 peer 10
run cfg1 peer 10
grep -h "Error\|error" $O/*.err | head -5

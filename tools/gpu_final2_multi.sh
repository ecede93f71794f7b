#!/bin/bash
# round 2, second set of final multi-GPU measurements (after the mid / warp / K5 kernel changes): cfg3, cfg4, cfg5 at N = $1 ranks
set -u
N=$1
O=gpurun_out/final2_n$N
mkdir -p $O
run() { # config steps
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --config $1 --gpus $N --steps $2 --warmup 3 --no-cpu --exchange peer > $O/bench_$1_n$N.json 2> $O/bench_$1_n$N.err
  echo "$1 n$N exit $?"; cut -c1-200 $O/bench_$1_n$N.json
}
run cfg3 10
run cfg4 5
run cfg5 5
grep -h "Error\|error" $O/*.err | head -5

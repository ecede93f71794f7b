#!/bin/bash
# round 2, final multi-GPU measurements: every BASELINE config at N = $1 ranks (peer-group exchange), cfg3 also with
# the NCCL collectives for comparison
set -u
N=$1
O=gpurun_out/final_n$N
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
run() { # config exchange steps suffix
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --config $1 --gpus $N --steps $3 --warmup 3 --no-cpu --exchange $2 > $O/bench_$1_n$N$4.json 2> $O/bench_$1_n$N$4.err
  echo "$1 n$N $2 exit $?"; cut -c1-200 $O/bench_$1_n$N$4.json
}
run cfg3 peer 10 ""
run cfg3 nccl 10 _nccl
run cfg1 peer 10 ""
run cfg2 peer 10 ""
run cfg4 peer 5 ""
run cfg5 peer 5 ""
grep -h "Error\|error" $O/*.err | head -5

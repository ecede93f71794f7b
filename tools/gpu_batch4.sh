#!/bin/bash
# round 2, batch 4 (2 GPUs): all -m gpu tests incl. the multi-device ones; cfg3 at N=1 (pipelined operator) and N=2
# (peer-group exchange vs NCCL collectives); cfg1 at N=2 (max exchange)
set -u
O=gpurun_out/b4
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
SSC_TRACE=1 timeout 600 python bench.py --config cfg3 --steps 5 --warmup 3 --no-cpu > $O/bench_cfg3_n1.json 2> $O/bench_cfg3_n1.err
echo "cfg3 n1 exit $?"; cut -c1-250 $O/bench_cfg3_n1.json
for ex in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --config cfg3 --gpus 2 --steps 10 --warmup 3 --exchange $ex > $O/bench_cfg3_n2_$ex.json 2> $O/bench_cfg3_n2_$ex.err
echo "cfg3 n2 $ex exit $?"; cut -c1-250 $O/bench_cfg3_n2_$ex.json
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --config cfg1 --gpus 2 --steps 10 --warmup 3 > $O/bench_cfg1_n2.json 2> $O/bench_cfg1_n2.err
echo "cfg1 n2 exit $?"; cut -c1-250 $O/bench_cfg1_n2.json
tail -3 $O/*.err

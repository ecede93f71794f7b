#!/bin/bash
# round 2, batch 17 (1 GPU): mid kernel with bounded merge loops; parity tests of the mid kernel; cfg4 slice at 4 / 3 resident blocks; cfg4 whole
set -u
O=gpurun_out/b17
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_shapes.py tests/test_gpu_device_layer.py -m gpu -x -q -k "mid or cfg4 or cfg5 or one_fma or median" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
for pad in 0 8192; do
echo "pad $pad: $(SSC_MID_SMEM_PAD=$pad python tools/profile_step.py --config cfg4 --rows 440 --steps 3 2>&1 | tail -1)"
done
timeout 300 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_cfg4.json 2> $O/bench_cfg4.err
python - <<PY
import json
d = json.loads(open("$O/bench_cfg4.json").read().strip().splitlines()[-1])
print("cfg4  ms/step %.3f" % d["ms_per_step"], d.get("step_breakdown_ms"), d.get("parity"))
PY

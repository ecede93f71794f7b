#!/bin/bash
# round 2, batch 21 (1 GPU): linear fit, literal pass out of line; parity tests; cfg5 whole and slice
set -u
O=gpurun_out/b21
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_shapes.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 300 python bench.py --config cfg5 --steps 3 --warmup 3 --no-cpu --no-e2e > $O/bench_cfg5.json 2> $O/bench_cfg5.err
python - <<PY
import json
d = json.loads(open("$O/bench_cfg5.json").read().strip().splitlines()[-1])
print("cfg5  ms/step %.3f" % d["ms_per_step"], d.get("step_breakdown_ms"), d.get("parity"))
PY
python tools/profile_step.py --config cfg5 --rows 152 --steps 3 2>&1 | tail -1

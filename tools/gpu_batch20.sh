#!/bin/bash
# round 2, batch 20 (1 GPU): linear fit with per-pass literal fallback inside the warp kernel (no hand-back chain); all -m gpu tests; cfg5 whole and slice
set -u
O=gpurun_out/b20
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 300 python bench.py --config cfg5 --steps 3 --warmup 3 --no-cpu --e2e-steps 2 > $O/bench_cfg5.json 2> $O/bench_cfg5.err
python - <<PY
import json
d = json.loads(open("$O/bench_cfg5.json").read().strip().splitlines()[-1])
print("cfg5  ms/step %.3f" % d["ms_per_step"], d.get("step_breakdown_ms"), d.get("parity"), "e2e", d.get("e2e", {}).get("ms_per_step"))
print(d["config"].get("rejection_counters"), d["config"].get("guard_fired"))
PY
python tools/profile_step.py --config cfg5 --rows 152 --steps 3 2>&1 | tail -2

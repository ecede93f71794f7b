#!/bin/bash
# round 2, batch 19 (1 GPU): K5 with shortcut searches and interleaved scans, mid kernel tail (integer winsorization walk, per-warp counters); all -m gpu tests
set -u
O=gpurun_out/b19
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
for c in cfg4 cfg3; do
timeout 300 python bench.py --config $c --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_$c.json 2> $O/bench_$c.err
echo "$c exit $?"; python - <<PY
import json
try:
    d = json.loads(open("$O/bench_$c.json").read().strip().splitlines()[-1])
    print("  ms/step %.3f" % d["ms_per_step"], d.get("step_breakdown_ms"), d.get("parity"))
except Exception as e:
    print("  no line:", e)
PY
done
python tools/profile_step.py --config cfg4 --rows 440 --steps 3 2>&1 | tail -2
python tools/profile_step.py --config cfg3 --rows 500 --steps 4 2>&1 | tail -2

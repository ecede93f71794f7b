#!/bin/bash
# round 2, batch 12 (1 GPU): the one-FMA normalization -- all -m gpu tests with the default build, then cfg3 (and cfg4)
# with the exact normalization, the default build (16-bit loads, SSC_CHEAP_VARIANT 2) and the conversion-based variant
set -u
O=gpurun_out/b12
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --config cfg3 --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_cfg3_$name.json 2> $O/bench_cfg3_$name.err
  echo "cfg3 $name exit $?"; python - <<PY
import json
try:
    d = json.loads(open("$O/bench_cfg3_$name.json").read().strip().splitlines()[-1])
    print("  ms/step %.3f" % d["ms_per_step"], d.get("step_breakdown_ms"), d.get("parity"))
except Exception as e:
    print("  no line:", e)
PY
}
run exact SSC_CHEAP_NORM=0
run v2 SSC_CHEAP_NORM=1
run v1 SSC_LIB=$PWD/build/variants/libssc_v1.so
run exact_again SSC_CHEAP_NORM=0
run v2_again SSC_CHEAP_NORM=1
SSC_LIB=$PWD/build/variants/libssc_v1.so timeout 600 python -m pytest tests/test_gpu_device_layer.py tests/test_gpu_parity.py -m gpu -x -q -k "one_fma or equals_general or cfg3" > $O/pytest_v1.log 2>&1; echo "pytest v1 exit $?"; tail -2 $O/pytest_v1.log
for e in 0 1; do
SSC_CHEAP_NORM=$e timeout 300 python bench.py --config cfg4 --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_cfg4_cheap$e.json 2> $O/bench_cfg4_cheap$e.err
echo "cfg4 cheap=$e exit $?"; cut -c1-220 $O/bench_cfg4_cheap$e.json
done

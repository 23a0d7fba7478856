#!/bin/bash
# full GPU suite + default bench (+ optional A/B env runs)
TAG=${1:-r02r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu --no-parity --no-e2e-run --e2e-steps 1 > $OUT/bench_quick.json 2> $OUT/bench_quick.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$OUT/bench_quick.json"))
print("ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["step_frac"], 4)); print({k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
PY
JPM_SIDE_STREAM=0 timeout 600 python bench.py --no-cpu --no-parity --no-e2e-run --e2e-steps 1 > $OUT/bench_noside.json 2> $OUT/bench_noside.err
python - <<PY
import json
d = json.load(open("$OUT/bench_noside.json"))
print("no side stream: ms/step", round(d["ms_per_step"], 4)); print({k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
PY

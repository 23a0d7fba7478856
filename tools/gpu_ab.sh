#!/bin/bash
# A/B bench runs at 512^3: each argument after the tag is "label:ENV1=v1,ENV2=v2:bench args"
# usage: gpurun -- bash tools/gpu_ab.sh r02c "spec::--force-mode spectral" "persist:JPM_READ_PERSIST=1024:--force-mode spectral"
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
if [ -n "$PYTEST_ARGS" ]; then
  timeout 1200 python -m pytest $PYTEST_ARGS -q ${PYTEST_K:+-k "$PYTEST_K"} > $OUT/pytest.log 2>&1; echo "== pytest rc=$?"; tail -n 12 $OUT/pytest.log
fi
for spec in "$@"; do
  label=${spec%%:*}; rest=${spec#*:}; envs=${rest%%:*}; args=${rest#*:}
  envs=${envs//,/ }
  env $envs timeout 900 python bench.py --no-cpu --no-e2e-run --e2e-steps 1 --steps 10 $args > $OUT/bench_$label.json 2> $OUT/bench_$label.err
  echo "== $label ($envs | $args) rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$label.json"))
    print("  ms/step", round(d["ms_per_step"], 4), "step_frac", round(d["roofline"]["step_frac"], 4), "force", d["force_path"]["steps_spectral"], d["force_path"]["steps_potential"], "bound", d["force_path"]["error_bound"], "parity", (d["parity"] or {}).get("final_pk_max_rel_diff"))
    print("  ", {k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("  no json:", e); print(open("$OUT/bench_$label.err").read()[-1200:])
PY
done

#!/bin/bash
# Round-2 first GPU call: the new tests (potential force path, fast functional API, golden slab fixtures, re-pinned
# config 1), then bench.py at 256^3 and 512^3 in the three force modes.  Outputs land in gpurun_out/<tag>/.
# usage: gpurun --timeout 1500 -- bash tools/gpu_r2a.sh [tag]
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/nvsmi.csv 2>&1
run() {  # name, timeout, command...
  local name=$1 to=$2; shift 2
  timeout $to "$@" > $OUT/$name.log 2>&1; local rc=$?
  echo "== $name rc=$rc"; tail -n ${TAILN:-6} $OUT/$name.log
}
run t_potential 600 python -m pytest tests/test_gpu_potential.py -q -s
run t_golden 600 python -m pytest tests/test_gpu_golden.py -q
run t_config1 900 python -m pytest tests/test_gpu_parity.py -q -k "config1 or sim_step or sim_load"
run t_slab 600 python -m pytest tests/test_gpu_slab.py -q
for mode in spectral potential auto; do
  timeout 600 python bench.py --size 256 --force-mode $mode --no-cpu --steps 10 > $OUT/bench256_$mode.json 2> $OUT/bench256_$mode.err
  echo "== bench256 $mode rc=$?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench256_$mode.json"))
    print("ms/step", round(d["ms_per_step"], 4), "force_path", d["force_path"], "parity", d["parity"], "api", d["api"], "lpt", d["lpt"])
    print({k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("no json:", e); print(open("$OUT/bench256_$mode.err").read()[-1500:])
PY
done
for mode in spectral potential auto; do
  timeout 900 python bench.py --force-mode $mode --no-cpu --steps 10 > $OUT/bench512_$mode.json 2> $OUT/bench512_$mode.err
  echo "== bench512 $mode rc=$?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench512_$mode.json"))
    print("ms/step", round(d["ms_per_step"], 4), "step_frac", round(d["roofline"]["step_frac"], 4), "force_path", d["force_path"])
    print("parity", d["parity"], "api", d["api"], "lpt", d["lpt"], "e2e", d["e2e"]["value"], "e2e_run", d["e2e_run"])
    print({k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("no json:", e); print(open("$OUT/bench512_$mode.err").read()[-1500:])
PY
done
TAILN=12 run t_scale 900 python -m pytest tests/test_gpu_scale.py -q -s -k 512
ls -la $OUT

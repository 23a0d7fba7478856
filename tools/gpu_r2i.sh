#!/bin/bash
# robustness: the full GPU suite twice + smoke + the default bench and the reference arm exactly as the driver runs them
TAG=${1:-r02u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for i in 1 2; do
  ( time timeout 1200 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu_$i.log 2>&1; echo "pytest $i rc=$?"; tail -4 $OUT/pytest_gpu_$i.log | cut -c1-200
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/smoke.log
( time timeout 900 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -4 $OUT/bench.err
( time timeout 900 python bench.py --impl reference ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "ref rc=$?"; tail -4 $OUT/bench_reference.err
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench.json") if l.startswith("{")][0])
print("ms/step", d["ms_per_step"], "frac", d["roofline"]["step_frac"], "e2e", d["e2e"]["value"], d["e2e"]["steps"], "e2e_run", d["e2e_run"]["value"], "cpu", d["cpu_baseline"]["value"], "parity", d["parity"]["final_pk_max_rel_diff"], "launches", d["gpu_launches"], d["clocks"])
r = json.loads([l for l in open("$OUT/bench_reference.json") if l.startswith("{")][0])
print("reference arm", r["value"], r["ms_per_step"], r["cpu_baseline"]["cores"])
PY

#!/bin/bash
# final round-2 evidence on one GPU: full suite, default bench, ncu --set full of one potential-chain step (exported)
TAG=${1:-r03e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 900 python -m pytest tests -m gpu -q ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest_gpu.log | cut -c1-200
( time timeout 600 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'sim_paint_kernel|sim_read_kernel|zinv_kernel|zfwd_kernel|yfwd_kernel|xpot_kernel|ypot_kernel|fdgrad_kernel' -s 330 -c 8 -f -o $OUT/prof_step_potential \
    python bench.py --no-cpu --no-parity --no-e2e-run --e2e-steps 1 --steps 3 --warmup 3 --force-mode potential > $OUT/full_run_potential.log 2>&1
echo "ncu rc=$?"
ncu -i $OUT/prof_step_potential.ncu-rep --page raw --csv > $OUT/prof_step_potential_raw.csv 2>/dev/null
rm -f $OUT/prof_step_potential.ncu-rep
python - <<PY
import json
d = json.loads([l for l in open("$OUT/bench.json") if l.startswith("{")][0])
print("ms/step", d["ms_per_step"], "frac", d["roofline"]["step_frac"], "e2e", d["e2e"]["value"], "parity", d["parity"]["final_pk_max_rel_diff"], d["clocks"])
print({k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
PY

#!/bin/bash
TAG=${1:-r03c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_potential.py tests/test_gpu_ic.py tests/test_gpu_slab.py -m gpu -q -x -k "fft or fused or potential or chain or linear_field or slab_forces or pencil_forces or sim_step" > $OUT/t_fft.log 2>&1; echo "fft tests rc=$?"; tail -3 $OUT/t_fft.log
timeout 300 python bench.py --no-cpu --no-e2e-run --e2e-steps 1 > $OUT/bench_quick.json 2> $OUT/bench_quick.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$OUT/bench_quick.json"))
print("ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["step_frac"], 4), d["parity"]["final_pk_max_rel_diff"], d["lpt"], d["api"]); print({k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
PY

#!/bin/bash
# quick bench + ncu evidence of the gradient (config 5) and order-preserving API kernels
TAG=${1:-r02s}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python bench.py --no-cpu --no-parity --no-e2e-run --e2e-steps 1 > $OUT/bench_quick.json 2> $OUT/bench_quick.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$OUT/bench_quick.json"))
print("ms/step", round(d["ms_per_step"], 4), "frac", round(d["roofline"]["step_frac"], 4)); print({k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/launches_grad.csv \
    python bench.py --grad --size 256 --steps 10 > $OUT/launches_grad_run.log 2>&1; echo "grad launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'readgrad_kernel|paintgrad_kernel|paint_direct_kernel|read_kernel|greens_div|kspace_kernel|pk_bin_kernel|pk_weight_kernel|lpt2_source|lpt2_shear' \
    -s 40 -c 12 -f -o $OUT/prof_grad python bench.py --grad --size 256 --steps 4 > $OUT/full_grad.log 2>&1; echo "grad full rc=$?"
ncu -i $OUT/prof_grad.ncu-rep --page raw --csv > $OUT/prof_grad_raw.csv 2>/dev/null
ls -la $OUT

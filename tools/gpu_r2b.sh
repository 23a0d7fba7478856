#!/bin/bash
# ncu --set full of the potential-chain kernels (one step late in the run) + the 768-thread variant of the read
TAG=${1:-r02b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on \
   -k regex:'sim_readpot_kernel|xpot_kernel|ypot_kernel|zinv_kernel' -s 152 -c 4 -f -o $OUT/prof_pot \
   python bench.py --no-cpu --no-parity --no-e2e-run --e2e-steps 1 --steps 3 --warmup 3 --force-mode potential > $OUT/full_run.log 2>&1
echo "ncu rc=$?"; tail -3 $OUT/full_run.log
JPM_POT_THREADS=768 timeout 600 python bench.py --force-mode potential --no-cpu --no-parity --no-e2e-run --e2e-steps 1 --steps 10 > $OUT/bench512_pot768.json 2> $OUT/bench512_pot768.err
python - <<PY
import json
d = json.load(open("$OUT/bench512_pot768.json"))
print("768 threads: ms/step", round(d["ms_per_step"], 4)); print({k: v["ms"] for k, v in d["roofline"]["kernels"].items()})
PY
for a in 0 2; do
JPM_L2_AHEAD=$a timeout 600 python bench.py --force-mode potential --no-cpu --no-parity --no-e2e-run --e2e-steps 1 --steps 10 > $OUT/bench512_l2a$a.json 2> $OUT/bench512_l2a$a.err
python - <<PY
import json
d = json.load(open("$OUT/bench512_l2a$a.json"))
print("l2ahead $a: ms/step", round(d["ms_per_step"], 4), d["roofline"]["kernels"]["tile_scan+sim_readpot_kick_drift"])
PY
done
ls -la $OUT

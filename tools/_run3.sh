#!/bin/bash
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python bench.py --no-cpu --e2e-steps 1 "$@" > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.load(open('$OUT/bench.json'))
print('ms_per_step',d['ms_per_step'],'step_frac',d['roofline']['step_frac'])
for k,v in d['roofline']['kernels'].items(): print(f"  {k:45s} {v['ms']:8.4f} ms {v['GBps']:8.1f} GB/s {v['frac']:.3f}")
print(d['sim'])
PY

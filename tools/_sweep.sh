#!/bin/bash
run() { echo "== $*"; env "$@" python bench.py --no-cpu --e2e-steps 1 --steps 5 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k[:28]:v['ms'] for k,v in d['roofline']['kernels'].items()})"; }
run JPM_PAINT_DEBUG=1
run JPM_PAINT_DEBUG=2

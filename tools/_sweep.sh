#!/bin/bash
run() { echo "== $*"; env "$@" python bench.py --no-cpu --e2e-steps 1 --steps 5 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k[:28]:v['ms'] for k,v in d['roofline']['kernels'].items() if 'read' in k or 'paint' in k})"; }
run JPM_L2_AHEAD=0
run JPM_L2_AHEAD=1
run JPM_L2_AHEAD=3
run JPM_L2_AHEAD=5

#!/bin/bash
# 8-GPU check, trimmed for box time: the (8,1) slab and (4,2) pencil parity cases, then bench at N = 8 and 4.
TAG=${1:-r01m8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 300 python -m pytest tests/test_multi_gpu.py -x -q -k "pdims5 or pdims3" > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 $OUT/pytest_multi.log
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --no-cpu > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
  echo "bench n=$n rc=$?"; grep "^{" $OUT/bench_n$n.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'ms/step', d['ms_per_step'], 'value', d['value'], d['timing'], 'e2e', d['e2e']['value'])"
  tail -2 $OUT/bench_n$n.err
done

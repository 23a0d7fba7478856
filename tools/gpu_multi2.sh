#!/bin/bash
# usage: gpurun --gpus N -- bash tools/gpu_multi2.sh <tag> <N list e.g. "2"> <modes e.g. "spectral auto"> [pytest -k expr]
TAG=$1; NS=$2; MODES=$3; KEXPR=$4
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
if [ -n "$KEXPR" ]; then
  timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -k "$KEXPR" > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -5 $OUT/pytest_multi.log
fi
for n in $NS; do
  for mode in $MODES; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --no-cpu --force-mode $mode $BENCH_ARGS > $OUT/bench_n${n}_$mode.json 2> $OUT/bench_n${n}_$mode.err
    echo "== bench n=$n $mode rc=$?"; grep "^{" $OUT/bench_n${n}_$mode.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', '%.3e' % d['value'], d['timing'], 'parity', d['parity'], 'force', d['force_path'], 'nvlink', d['nvlink'])"
    tail -2 $OUT/bench_n${n}_$mode.err
  done
done

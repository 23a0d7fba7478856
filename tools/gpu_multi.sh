#!/bin/bash
# Multi-GPU check: usage: gpurun --gpus N --timeout 1500 -- bash tools/gpu_multi.sh <tag> <N> [extra bench args]
TAG=${1:-r01m}; N=${2:-2}; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/nvsmi.csv 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -4 $OUT/pytest_multi.log
for n in $(seq 1 $N); do
  if [[ $n == 1 || $n == 2 || $n == 4 || $n == 8 ]]; then
    if [[ $n == 1 ]]; then
      timeout 600 python bench.py --no-cpu "$@" > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
          bench.py --gpus $n --no-cpu "$@" > $OUT/bench_n$n.json 2> $OUT/bench_n$n.err
    fi
    echo "bench n=$n rc=$?"; tail -c 1500 $OUT/bench_n$n.json; tail -3 $OUT/bench_n$n.err
  fi
done

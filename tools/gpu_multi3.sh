#!/bin/bash
# Multi-GPU check incl. pencil grids: usage: gpurun --gpus N --timeout 1500 -- bash tools/gpu_multi3.sh <tag> <N>
TAG=${1:-r02m}; N=${2:-4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -q > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -6 $OUT/pytest_multi.log
run() {  # n, tag, extra args
  local n=$1 tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --no-cpu "$@" > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  echo "bench $tag rc=$?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_$tag.json; grep -o '"parity": {[^}]*}' $OUT/bench_$tag.json | cut -c1-160
  grep -o '"stage_ms_max_over_ranks": {[^}]*}' $OUT/bench_$tag.json; tail -2 $OUT/bench_$tag.err | cut -c1-300
}
if [[ $N == 4 ]]; then
  run 4 n4_slab
  run 4 n4_2x2 --pdims 2x2
  run 4 n4_1x4 --pdims 1x4 --no-parity
  run 2 n2_slab --no-parity
  run 2 n2_1x2 --pdims 1x2 --no-parity
fi
if [[ $N == 8 ]]; then
  run 8 n8_slab
  run 8 n8_2x4 --pdims 2x4
  run 8 n8_4x2 --pdims 4x2 --no-parity
  run 8 n8_2x4_nccl --pdims 2x4 --nccl --no-parity
  run 4 n4_slab --no-parity
  run 2 n2_slab --no-parity
fi

#!/bin/bash
# Final 8-GPU evidence: multi-process parity of the 8-rank grids, strong-scaling bench (slabs, pencil grids fused and
# over NCCL), BASELINE.json config 4 (1024^3).   usage: gpurun --gpus 8 --timeout 1500 -- bash tools/gpu_multi8.sh <tag>
TAG=${1:-r02m8b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -q -k "pdims3 or pdims4 or pdims5" > $OUT/pytest_multi.log 2>&1; echo "pytest multi rc=$?"
tail -4 $OUT/pytest_multi.log
run() {  # tag, extra args
  local tag=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 8 --no-cpu "$@" > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  echo "bench $tag rc=$?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_$tag.json; grep -o '"parity": {[^}]*}' $OUT/bench_$tag.json | cut -c1-160
  grep -o '"stage_ms_max_over_ranks": {[^}]*}' $OUT/bench_$tag.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\*" $OUT/bench_$tag.err | tail -3 | cut -c1-300
}
run n8_slab
run n8_2x4 --pdims 2x4
run n8_4x2 --pdims 4x2 --no-parity
run n8_2x4_nccl --pdims 2x4 --nccl --no-parity --steps 5
run n8_1024 --size 1024 --no-parity --steps 5 --warmup 3 --e2e-steps 1
nvidia-smi --query-gpu=index,memory.used --format=csv | head -3

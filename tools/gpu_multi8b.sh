#!/bin/bash
# 8 GPUs: slabs (8x1) against 4x2 pencils, twice each, on the same box
TAG=${1:-r02m8c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {
  local tag=$1; shift
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 8 --no-cpu --no-parity --e2e-steps 1 "$@" > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  echo "bench $tag rc=$?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_$tag.json
  grep -o '"stage_ms_max_over_ranks": {[^}]*}' $OUT/bench_$tag.json; grep -o '"stage_ms_min_over_ranks": {[^}]*}' $OUT/bench_$tag.json
}
run slab_a
run p4x2_a --pdims 4x2
run slab_b
run p4x2_b --pdims 4x2

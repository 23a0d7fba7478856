#!/bin/bash
# what does the ghost exchange cost?  N-GPU slab bench with the ghost width forced (timing only; physics wrong for 0)
TAG=${1:-r02m}; N=${2:-4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() {
  local tag=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --no-cpu --no-parity --force-mode potential > $OUT/bench_$tag.json 2> $OUT/bench_$tag.err
  echo "bench $tag rc=$?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_$tag.json
  grep -o '"stage_ms_max_over_ranks": {[^}]*}' $OUT/bench_$tag.json; grep -o '"ghost_planes_used": [0-9]*' $OUT/bench_$tag.json
}
run normal JPM_X=0
run ghost0 JPM_SLAB_GHOST_OVERRIDE=0
run ghost8 JPM_SLAB_GHOST_OVERRIDE=8
run ghost16 JPM_SLAB_GHOST_OVERRIDE=16

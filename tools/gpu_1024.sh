#!/bin/bash
# config 4 of BASELINE.json: 1024^3 particles on a 1024^3 mesh over 8 GPUs (slab path), short timed window
OUT=gpurun_out/${1:-r01g1024}; mkdir -p $OUT
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 8 --size 1024 --no-cpu --steps 5 --warmup 3 --e2e-steps 1 > $OUT/bench_1024_n8.json 2> $OUT/bench_1024_n8.err
echo "rc=$?"; grep "^{" $OUT/bench_1024_n8.json | cut -c1-3000; tail -5 $OUT/bench_1024_n8.err
nvidia-smi --query-gpu=index,memory.used --format=csv | head -3

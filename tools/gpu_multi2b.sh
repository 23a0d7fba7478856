#!/bin/bash
# 2 GPUs: the slab bench line with the driver's NVLink counters next to the byte model
TAG=${1:-r03i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi nvlink -gt d -i 0 | head -8 > $OUT/nvlink_sample.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --no-cpu --no-parity --e2e-steps 1 > $OUT/bench_n2.json 2> $OUT/bench_n2.err
echo "rc=$?"; grep -o '"ms_per_step": [0-9.]*' $OUT/bench_n2.json; grep -o '"nvlink": {.*}' $OUT/bench_n2.json | cut -c1-700; head -5 $OUT/nvlink_sample.txt

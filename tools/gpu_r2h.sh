#!/bin/bash
TAG=${1:-r02t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_gradients.py tests/test_gpu_transforms.py -m gpu -q -x > $OUT/t_grad.log 2>&1; echo "grad tests rc=$?"; tail -15 $OUT/t_grad.log | cut -c1-300
timeout 600 python bench.py --grad --size 256 --steps 10 > $OUT/bench_grad256.json 2> $OUT/bench_grad256.err; echo "grad rc=$?"; cat $OUT/bench_grad256.json | cut -c1-900
JPM_FUSED_VJP=0 timeout 600 python bench.py --grad --size 256 --steps 10 > $OUT/bench_grad256_unfused.json 2> $OUT/bench_grad256_unfused.err; echo "grad unfused rc=$?"; cat $OUT/bench_grad256_unfused.json | cut -c1-400

#!/bin/bash
mkdir -p gpurun_out/dbg
P=tools/tma_probe
{
for args in "1 16 11 11 0 3 -1 32" "3 16 11 11 0 3 -1 32" "1 16 11 11 0 3 3 32" "3 16 11 11 0 3 3 32" "1 16 11 11 24 27 23 32" "3 16 11 11 24 27 23 32" "1 24 19 19 0 3 -1 64" "3 24 19 19 0 3 -1 64"; do
  timeout 30 $P $args
done
} 2>&1 | tee gpurun_out/dbg/tma_probe.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_slab.py -x -q -k "stepper and shape0" 2>&1 | grep -v "^$" | head -60 > gpurun_out/dbg/sanitizer.log
head -40 gpurun_out/dbg/sanitizer.log

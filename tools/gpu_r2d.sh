#!/bin/bash
# config-2 parity test with its log, the new pipelined e2e entry, then the ncu evidence (reports exported on the box)
TAG=${1:-r02h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -s -k config2 > $OUT/t_config2.log 2>&1; echo "config2 rc=$?"; tail -30 $OUT/t_config2.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or host_step" > $OUT/t_pipe.log 2>&1; echo "pipe rc=$?"; tail -5 $OUT/t_pipe.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; cat $OUT/bench.json; tail -5 $OUT/bench.err
bash tools/gpu_r2c.sh $TAG launches,full

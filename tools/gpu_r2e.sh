#!/bin/bash
TAG=${1:-r02i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -s -k config2 > $OUT/t_config2.log 2>&1; echo "config2 rc=$?"; grep "config2\|passed\|failed\|Error" $OUT/t_config2.log | cut -c1-400 | tail -30

#!/bin/bash
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_fft_chain or sim_step" > $OUT/pytest_fft.log 2>&1; echo "pytest fft rc=$?"; tail -3 $OUT/pytest_fft.log
bash tools/_run3.sh $TAG "$@"

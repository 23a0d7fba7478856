#!/bin/bash
# one pytest selection on one GPU: bash tools/gpu_quick.sh <tag> <pytest args...>
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest "$@" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest.log | cut -c1-300

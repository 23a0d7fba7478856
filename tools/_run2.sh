#!/bin/bash
TAG=$1
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "fused_fft_chain or sim_step" > $OUT/pytest_fft.log 2>&1; echo "pytest fft rc=$?"; tail -25 $OUT/pytest_fft.log
timeout 600 python bench.py --no-cpu --e2e-steps 1 --margin 1 > $OUT/bench_m1.json 2> $OUT/bench_m1.err; echo "bench m1 rc=$?"; cat $OUT/bench_m1.json; tail -3 $OUT/bench_m1.err
JPM_PMFFT=0 timeout 600 python bench.py --no-cpu --e2e-steps 1 --margin 1 > $OUT/bench_m1_cufft.json 2> $OUT/bench_m1_cufft.err; echo "bench cufft rc=$?"; python -c "
import json;d=json.load(open('$OUT/bench_m1_cufft.json'));print('cufft ms_per_step',d['ms_per_step'])"

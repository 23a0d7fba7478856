#!/bin/bash
TAG=$1
OUT=gpurun_out/$TAG; mkdir -p $OUT
bash tools/tma_probe.sh > /dev/null 2>&1; cat gpurun_out/dbg/tma_probe.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu --e2e-steps 1 --margin 1 > $OUT/bench_m1.json 2> $OUT/bench_m1.err; echo "bench m1 rc=$?"; cat $OUT/bench_m1.json; tail -3 $OUT/bench_m1.err
timeout 600 python bench.py --no-cpu --e2e-steps 1 --margin 2 > $OUT/bench_m2.json 2> $OUT/bench_m2.err; echo "bench m2 rc=$?"; cat $OUT/bench_m2.json; tail -3 $OUT/bench_m2.err

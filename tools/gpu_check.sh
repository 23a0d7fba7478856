#!/bin/bash
# One gpurun call: GPU parity tests, smoke, the 1-GPU bench, the ncu launch list and one
# `--set full` capture of the library's own hot kernels.  Outputs land in gpurun_out/.
# usage: gpurun --timeout 1500 -- bash tools/gpu_check.sh [tag] [stages]
TAG=${1:-r01}
STAGES=${2:-test,smoke,bench,launches,full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/nvsmi.csv 2>&1
if [[ $STAGES == *test* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
if [[ $STAGES == *smoke* ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
  tail -3 $OUT/smoke.log
fi
if [[ $STAGES == *bench* ]]; then
  nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $OUT/clocks.csv &
  SMI=$!
  timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
  kill $SMI
  cat $OUT/bench.json; tail -5 $OUT/bench.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?"
  cat $OUT/bench_reference.json
fi
if [[ $STAGES == *launches* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
      --log-file $OUT/launches.csv python bench.py --no-cpu --e2e-steps 1 > $OUT/launches_run.log 2>&1
  echo "launches rc=$?"
fi
if [[ $STAGES == *full* ]]; then
  # one `ncu --set full` capture of every own kernel of one PM step (the 38th: 37 untimed steps before it)
  timeout 1200 ncu --set full --clock-control none --import-source on \
      -k regex:'sim_paint_kernel|sim_read_kernel|xfused_kernel|zinv_kernel|zfwd_kernel|yfwd_kernel|yinv_kernel' -s 266 -c 7 -f -o $OUT/prof_step \
      python bench.py --no-cpu --e2e-steps 1 --steps 3 --warmup 3 > $OUT/full_run.log 2>&1
  echo "full rc=$?"
fi
ls -la $OUT

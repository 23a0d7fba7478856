#!/bin/bash
# Round-2 evidence run on one GPU: full GPU test suite, smoke, default bench + reference arm, config-5 gradient bench,
# ncu launch list, and `--set full` captures of the step kernels in both force modes.
TAG=${1:-r02g}
STAGES=${2:-test,smoke,bench,grad,launches,full}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/nvsmi.csv 2>&1
if [[ $STAGES == *test* ]]; then
  ( time timeout 1200 python -m pytest tests -m gpu -q --durations=15 ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
  tail -8 $OUT/pytest_gpu.log
fi
if [[ $STAGES == *smoke* ]]; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
  tail -3 $OUT/smoke.log
fi
if [[ $STAGES == *bench* ]]; then
  ( time timeout 900 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
  cat $OUT/bench.json; tail -5 $OUT/bench.err
  ( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?"
  cat $OUT/bench_reference.json; tail -4 $OUT/bench_reference.err
fi
if [[ $STAGES == *grad* ]]; then
  timeout 600 python bench.py --grad --size 256 --steps 10 > $OUT/bench_grad256.json 2> $OUT/bench_grad256.err; echo "grad rc=$?"
  cat $OUT/bench_grad256.json; tail -5 $OUT/bench_grad256.err
fi
if [[ $STAGES == *launches* ]]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
      --log-file $OUT/launches.csv python bench.py --no-cpu --e2e-steps 1 > $OUT/launches_run.log 2>&1
  echo "launches rc=$?"
fi
if [[ $STAGES == *full* ]]; then
  for mode in potential spectral; do
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:'sim_paint_kernel|sim_read_kernel|xfused_kernel|zinv_kernel|zfwd_kernel|yfwd_kernel|yinv_kernel|xpot_kernel|ypot_kernel|fdgrad_kernel' -s 330 -c 8 -f -o $OUT/prof_step_$mode \
      python bench.py --no-cpu --no-parity --no-e2e-run --e2e-steps 1 --steps 3 --warmup 3 --force-mode $mode > $OUT/full_run_$mode.log 2>&1
  echo "full $mode rc=$?"
  # gpurun brings back <= 64 MiB: export what the summaries need, keep the report only for the potential chain
  ncu -i $OUT/prof_step_$mode.ncu-rep --page raw --csv > $OUT/prof_step_${mode}_raw.csv 2>/dev/null
  ncu -i $OUT/prof_step_$mode.ncu-rep --page source --csv --print-source cuda,sass > $OUT/prof_step_${mode}_src.csv 2>/dev/null
  gzip -f $OUT/prof_step_${mode}_src.csv
  [[ $mode == spectral ]] && rm -f $OUT/prof_step_$mode.ncu-rep
  done
fi
ls -la $OUT

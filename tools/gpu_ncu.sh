#!/bin/bash
# ncu --set full of every kernel of one late PM step (auto force mode -> potential chain), plus the launch list
# usage: gpurun -- bash tools/gpu_ncu.sh <tag> [force-mode]
TAG=${1:-r02n}; MODE=${2:-auto}
OUT=gpurun_out/$TAG; mkdir -p $OUT
REGEX='sim_paint_kernel|sim_read_kernel|xpot_kernel|ypot_kernel|zinv_kernel|zfwd_kernel|yfwd_kernel|fdgrad_kernel|xfused_kernel|yinv_kernel'
# kernels matching before step 38: LPT (zfwd/yfwd/xfused/yinv/zinv per force evaluation) + 37 steps; skip generously
# and take one full step's worth (8 kernels in potential mode, 7 in spectral)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s ${SKIP:-330} -c ${CNT:-9} -f -o $OUT/prof_step \
   python bench.py --no-cpu --no-parity --no-e2e-run --e2e-steps 1 --steps 3 --warmup 3 --force-mode $MODE > $OUT/full_run.log 2>&1
echo "ncu full rc=$?"; tail -2 $OUT/full_run.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/launches.csv \
   python bench.py --no-cpu --no-parity --no-e2e-run --e2e-steps 1 --steps 3 --warmup 3 --force-mode $MODE > $OUT/launches_run.log 2>&1
echo "launches rc=$?"
ls -la $OUT

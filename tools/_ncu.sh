#!/bin/bash
# one `ncu --set full` capture of every own kernel of one PM step (after 35 untimed steps)
TAG=$1; SKIP=${2:-266}; CNT=${3:-7}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 ncu --set full --clock-control none --import-source on \
   -k regex:'sim_paint_kernel|sim_read_kernel|xfused_kernel|zinv_kernel|zfwd_kernel|yfwd_kernel|yinv_kernel' -s $SKIP -c $CNT -f -o $OUT/prof_step \
   python bench.py --no-cpu --e2e-steps 1 --steps 3 --warmup 3 > $OUT/full_run.log 2>&1
echo "ncu rc=$?"; tail -5 $OUT/full_run.log; ls -la $OUT

#!/bin/bash
# runs the TMA probe variants, each in its own process (an illegal instruction kills the context)
P=tools/tma_probe
mkdir -p gpurun_out/dbg
{
for args in \
 "0 16 13 13 4 2 2 40" "0 24 21 21 -4 -2 -2 64" "0 24 21 21 48 46 46 64" "3 24 21 21 16 14 14 64" "3 24 21 21 -4 30 46 64" \
 "1 24 21 21 16 14 14 64" "1 24 21 21 -4 -2 -2 64" "1 24 21 21 48 46 46 64" ; do
  timeout 30 $P $args
done
} 2>&1 | tee gpurun_out/dbg/tma_probe.log

#!/bin/bash
# GPU run 5 of round 2: pencil tile order (SSAM_TILE_ORDER=2) with the percentile prefetch distance, v1 kernel
mkdir -p gpurun_out
out=gpurun_out/r2_pencil.jsonl; : > $out
run() { timeout 300 python profiles/experiments/r2_one.py "$@" >> $out 2>>gpurun_out/r2_run5.err; }
for wl in cfg2 cfg4 cfg5q cfg5y; do
  SSAM_TILE_ORDER=0 run $wl 1
  SSAM_TILE_ORDER=2 run $wl 1
  SSAM_TILE_ORDER=1 run $wl 1
done
SSAM_TILE_ORDER=2 run cfg2 2
SSAM_TILE_ORDER=2 run cfg5q 2
cat $out

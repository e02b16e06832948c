#!/bin/bash
# GPU run 4 of round 2: tile traversal order on 512-wide planes (cfg5q 512x512x64, cfg5y 256x512x128)
mkdir -p gpurun_out
out=gpurun_out/r2_tile_order.jsonl; : > $out
run() { timeout 300 python profiles/experiments/r2_one.py "$@" >> $out 2>>gpurun_out/r2_run4.err; }
for wl in cfg5q cfg5y; do
  SSAM_TILE_ORDER=0 SSAM_BFS_ORDER=1 run $wl 1
  SSAM_TILE_ORDER=0 SSAM_BFS_ORDER=0 run $wl 1
  SSAM_TILE_ORDER=0 SSAM_BFS_ORDER=0 SSAM_PF=0 run $wl 1
  SSAM_TILE_ORDER=1 SSAM_BFS_ORDER=0 run $wl 1
  SSAM_TILE_ORDER=1 SSAM_BFS_ORDER=0 SSAM_PF=0 run $wl 1
  SSAM_TILE_ORDER=0 SSAM_BFS_ORDER=1 run $wl 2
  SSAM_TILE_ORDER=0 SSAM_BFS_ORDER=0 run $wl 2
  SSAM_TILE_ORDER=1 SSAM_BFS_ORDER=0 run $wl 2
done
cat $out

#!/bin/bash
# GPU run 3 of round 2: why is cfg5 (SheppardWright, 512x512 planes) at 0.68 when cfg4 reaches 0.80?
export SSAM_TILE_ORDER=0
mkdir -p gpurun_out
out=gpurun_out/r2_cfg5_matrix.jsonl; : > $out
for wl in cfg2 cfg2sw cfg5q cfg5qfks cfg5qnh cfg5y cfg4; do
  timeout 300 python profiles/experiments/r2_one.py $wl 1 >> $out 2>>gpurun_out/r2_run3.err
done
SSAM_LAYOUT=0 timeout 300 python profiles/experiments/r2_one.py cfg5q 1 >> $out 2>>gpurun_out/r2_run3.err
cat $out
timeout 300 python profiles/experiments/r2_one.py cfg5q 1 3 > gpurun_out/r2_plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cells_tiled -s 3 -c 1 -o gpurun_out/r2_tiled_cfg5q python profiles/experiments/r2_one.py cfg5q 1 3 > gpurun_out/r2_ncu3.log 2>&1
echo ncu_exit=$?; tail -3 gpurun_out/r2_ncu3.log
timeout 600 python -m pytest tests/test_gpu_conditioning.py -x -q -s > gpurun_out/r2_conditioning.log 2>&1; echo cond_exit=$?; tail -30 gpurun_out/r2_conditioning.log

#!/bin/bash
# GPU run 2 of round 2: parity on every variant, the persistent kernel's scheduling switches, one ncu capture
export SSAM_TILE_ORDER=0
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest2.log 2>&1; echo pytest_exit=$?; tail -3 gpurun_out/r2_pytest2.log
out=gpurun_out/r2_pipe_switches.jsonl; : > $out
for wl in cfg2 cfg5q; do
  timeout 300 python profiles/experiments/r2_one.py $wl 1 >> $out 2>>gpurun_out/r2_run2.err
  for dyn in 1 0; do for st in 0 3000 6000 12000; do
    SSAM_PIPE_DYNAMIC=$dyn SSAM_PIPE_STAGGER_NS=$st timeout 300 python profiles/experiments/r2_one.py $wl 2 >> $out 2>>gpurun_out/r2_run2.err
  done; done
  SSAM_PF_PIPE=148 timeout 300 python profiles/experiments/r2_one.py $wl 2 >> $out 2>>gpurun_out/r2_run2.err
done
cat $out
timeout 300 python profiles/experiments/r2_one.py cfg2 2 3 > gpurun_out/r2_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cells_pipe -s 3 -c 1 -o gpurun_out/r2_pipe_cfg2 python profiles/experiments/r2_one.py cfg2 2 3 > gpurun_out/r2_ncu.log 2>&1
echo ncu_exit=$?; tail -3 gpurun_out/r2_ncu.log

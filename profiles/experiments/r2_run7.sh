#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_local_ranks.py tests/test_openfoam_shim.py -x -q -s > gpurun_out/r2_pytest7a.log 2>&1; echo pytest7a_exit=$?; tail -25 gpurun_out/r2_pytest7a.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "alpha_compression or layout" > gpurun_out/r2_pytest7b.log 2>&1; echo pytest7b_exit=$?; tail -8 gpurun_out/r2_pytest7b.log
out=gpurun_out/r2_fits_ab.jsonl; : > $out
for wl in cfg2 cfg5q; do
  SSAM_FITS=1 timeout 300 python profiles/experiments/r2_one.py $wl 1 >> $out 2>>gpurun_out/r2_run7.err
  SSAM_FITS=0 timeout 300 python profiles/experiments/r2_one.py $wl 1 >> $out 2>>gpurun_out/r2_run7.err
done
cat $out

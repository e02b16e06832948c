#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_local_ranks.py -x -q -s > gpurun_out/r2_pytest8a.log 2>&1; echo pytest8a_exit=$?; tail -12 gpurun_out/r2_pytest8a.log
timeout 400 python -m pytest tests/test_openfoam_shim.py -x -q -s > gpurun_out/r2_pytest8b.log 2>&1; echo pytest8b_exit=$?; tail -12 gpurun_out/r2_pytest8b.log
out=gpurun_out/r2_fits_ab.jsonl; : > $out
for wl in cfg2 cfg5q; do
  SSAM_FITS=1 timeout 300 python profiles/experiments/r2_one.py $wl 1 >> $out 2>>gpurun_out/r2_run8.err
  SSAM_FITS=0 timeout 300 python profiles/experiments/r2_one.py $wl 1 >> $out 2>>gpurun_out/r2_run8.err
done
cat $out

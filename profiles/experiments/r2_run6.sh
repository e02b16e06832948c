#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest6.log 2>&1; echo pytest_exit=$?; tail -15 gpurun_out/r2_pytest6.log
( /usr/bin/env time -v true ) 2>/dev/null
start=$(date +%s)
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo bench_exit=$? elapsed=$(( $(date +%s) - start ))s
cat gpurun_out/r2_bench_n1.json; tail -25 gpurun_out/r2_bench_n1.err

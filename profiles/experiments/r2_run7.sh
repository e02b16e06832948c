#!/bin/bash
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_local_ranks.py tests/test_openfoam_shim.py -x -q -s > gpurun_out/r2_pytest7a.log 2>&1; echo pytest7a_exit=$?; tail -25 gpurun_out/r2_pytest7a.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "alpha_compression or layout" > gpurun_out/r2_pytest7b.log 2>&1; echo pytest7b_exit=$?; tail -8 gpurun_out/r2_pytest7b.log

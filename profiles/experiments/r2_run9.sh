#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest9.log 2>&1; echo pytest9_exit=$?; tail -15 gpurun_out/r2_pytest9.log

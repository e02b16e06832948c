#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r2_pytest10.log 2>&1; echo pytest10_exit=$?; tail -12 gpurun_out/r2_pytest10.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo bench2_exit=$?
cat gpurun_out/r2_bench_n2.json | cut -c1-3000; grep -v "^W\|^\[W\|warn" gpurun_out/r2_bench_n2.err | tail -15
SSAM_HALO=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 --no-parity > gpurun_out/r2_bench_n2_nccl.json 2> gpurun_out/r2_bench_n2_nccl.err; echo bench2nccl_exit=$?
cat gpurun_out/r2_bench_n2_nccl.json | cut -c1-700

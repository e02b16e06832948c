#!/bin/bash
# 8 GPUs: strong scaling of the 64M-cell box, 2x2x2 blocks and z-slabs, peer-memory and NCCL transports
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # name, extra env/args...
  name=$1; shift
  env "$@" timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29540 + RANDOM % 50)) bench.py --gpus 8 --steps 20 --warmup 5 $ARGS > gpurun_out/r2_bench_n8_$name.json 2> gpurun_out/r2_bench_n8_$name.err
  echo "$name exit=$?"
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_n8_$name.json") if l.startswith('{')][-1])
    print("$name", d["value"], d["ms_per_step"], "sustained", d["sustained"]["value"], "kernel", d["roofline"]["kernel_ms"], d["roofline"]["frac"], "e2e", d["e2e"]["value"], d["config"]["parallelism"], d["detail"]["halo_transport"], d["detail"]["host_numa_binding"], d["parity"] and d["parity"]["max_rel"])
except Exception as e:
    print("$name failed:", e)
PY
  grep -v "^W\|warn\|^\*\|OMP_NUM" gpurun_out/r2_bench_n8_$name.err | tail -4
}
ARGS="" run block_peer A=1
ARGS="--decomp slab" run slab_peer A=1
ARGS="--no-parity" run block_nccl SSAM_HALO=nccl

#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_local_ranks.py tests/test_multigpu.py -x -q > gpurun_out/r2_pytest11.log 2>&1; echo pytest11_exit=$?; tail -6 gpurun_out/r2_pytest11.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; echo bench2_exit=$?
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r2_bench_n2b.json") if l.startswith('{')][-1])
print(d["value"], d["ms_per_step"], d["sustained"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["kernel_share_of_step"], d["e2e"]["value"], d["parity"]["max_rel"])
PY

#!/bin/bash
# Final profile of round 2 on the headline workload (cfg5, 67.1M cells, one B200): plain run, launch list, one
# ncu --set full capture of the dominant kernel.  Each ncu pass follows the same command having exited 0 without ncu.
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-parity"
timeout 400 $CMD > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err || { echo plain-run-failed; tail -5 gpurun_out/r2_prof_plain.err; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg5.csv $CMD > gpurun_out/r2_ncu1.log 2>&1; echo launchlist_exit=$?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cells_tiled -s 3 -c 1 -f -o gpurun_out/r2_k_cells_cfg5 $CMD > gpurun_out/r2_ncu2.log 2>&1; echo full_exit=$?
ls -la gpurun_out/r2_k_cells_cfg5.ncu-rep gpurun_out/r2_launches_cfg5.csv

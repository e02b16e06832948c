set -x
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/prof_plain.json 2>/dev/null || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_compact.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cells_tiled -s 3 -c 1 -f -o gpurun_out/k_cells_compact python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
python - <<'PY'
import time, numpy as np
from solidstateamfoam_b200 import capi, cases, params as P
case = cases.cfg2()
for mode in (0, 1):
    c = capi.Context(0); c.set_cell_layout(mode)
    t = time.time(); c.mesh_set(case.mesh, case.patch_u_bc); dt = time.time() - t
    print("mesh_set layout", mode, "compact" if c.layout_is_compact() else "caller", round(dt, 2), "s")
    c.close()
PY

#!/usr/bin/env python
"""bench.py -- Mcell-updates/s of the fused grad(U) -> nu -> mixture -> heat-source update.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg4|cfg5|cfg3] [--impl reference]

Contract (one JSON line on stdout from rank 0):
  * a "step" = one fused update (ssam_update_fused) over the whole mesh of the workload;
  * N=1 workload = BASELINE.json configs[1]: blockMesh hex 256x256x128 (8.4M cells), FKS + cubic k(T);
    N>1 = weak scaling: every rank owns a 256x256x128 z-slab of a 256x256x(128N) box, processor-patch
    halos of U, temp, alpha1 (inputs) and epsilonDotEq (outputs) exchanged with NCCL every step;
  * value = cells of all ranks * K / max-over-ranks device time (CUDA events on the launching stream,
    barrier + synchronize on both sides), inputs resident in HBM;
  * roofline = the dominant kernel (k_cells) timed live with its own CUDA events inside the library
    (ssam_profile_*), algorithmic bytes = 40*N_if + 52*N_bf + 128*N_c;
  * e2e = the same update through ssam_update_fused_host with pinned HOST buffers: H2D of U, temp,
    alpha1, p and D2H of every output field inside the timed region;
  * cpu_baseline / --impl reference = the oracle's faithful field-at-a-time restatement of the
    reference (OpenFOAM itself is unavailable offline), one z-slab partition per host thread.
"""
from __future__ import annotations

import os as _os

# 5 MB halo messages: 4 NCCL channels are enough and leave more SMs to the interior tiles (profiles/r1_scaling.md)
_os.environ.setdefault("NCCL_MAX_NCHANNELS", "4")
import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Mcell-updates/s for grad(U)->nu->heat-source"
UNIT = "Mcell-updates/s"

WORKLOADS = {
    # name: (dims per rank, description)
    "cfg2": ((256, 256, 128), "blockMesh hex 256x256x128 (8.4M cells), FKS + cubic k(T), constantSlipStick"),
    "cfg4": ((256, 256, 256), "hex 256^3 fluid region (16.8M cells), NortonHoff, qfric skipped"),
    "cfg5": ((512, 512, 32), "z-slab of the 512x512x(32N) AFSD box, SheppardWright"),
    "cfg3": ((128, 128, 128), "polyhedral ~4M cells (2 levels of 2:1 refinement), SheppardWright, exponentialSlipStick"),
    "tiny": ((32, 32, 16), "debug"),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_case(workload: str, rank: int, n_ranks: int, dims=None):
    from solidstateamfoam_b200 import cases

    dims = dims or WORKLOADS[workload][0]
    nx, ny, nz = dims
    if workload in ("cfg2", "tiny"):
        cell = 0.25e-3
        side = [-1] * 6
        if rank > 0:
            side[4] = rank - 1
        if rank < n_ranks - 1:
            side[5] = rank + 1
        box = ((0.0, 0.0, 0.0), (nx * cell, ny * cell, n_ranks * nz * cell))
        return cases.cfg2(dims=dims, side=tuple(side), rank=rank, origin=(0.0, 0.0, rank * nz * cell),
                          z0=0.6 * n_ranks * nz * cell, box=box)
    if workload == "cfg5":
        return cases.cfg5_slab(rank, n_ranks, dims_per_rank=dims)
    if n_ranks != 1:
        raise SystemExit(f"workload {workload} is single-GPU only")
    if workload == "cfg4":
        return cases.cfg4(dims=dims)
    if workload == "cfg3":
        return cases.cfg3(base=dims)
    raise SystemExit(f"unknown workload {workload}")


# ---- clocks ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [x.strip() for x in s.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU arm -------------------------------------------------------------------------------------------
def cpu_arm(workload: str, steps: int, warmup: int, threads: int | None = None, layers_per_thread: int | None = None):
    """The oracle's faithful field-at-a-time update (pass structure of the reference), one z-slab
    partition per host thread with in-memory halo copies (one MPI rank per core).  Returns
    (Mcell/s, threads, sample description, ms_per_step)."""
    from oracle import pyoracle as O
    from solidstateamfoam_b200 import params as P

    threads = threads or os.cpu_count() or 1
    nx, ny, nz = WORKLOADS[workload][0]
    if workload == "cfg3":
        nx = ny = nz = 48
    lpt = layers_per_thread or max(1, min(nz // threads if nz >= threads else 1, 8))
    wl = "cfg2" if workload in ("cfg3", "cfg4") else workload  # hex slabs of the same lateral size
    cs = [build_case(wl, r, threads, dims=(nx, ny, lpt)) for r in range(threads)]
    if workload in ("cfg4", "cfg3"):  # same slabs, that workload's constitutive models
        src = build_case(workload, 0, 1, dims=(8, 8, 8))
        for c in cs:
            c.params.visc1 = src.params.visc1
            c.params.mix = src.params.mix
            if workload == "cfg4":
                c.params.stick.fricPatch = -1
    ocs = []
    for c in cs:
        oc = O.OracleCase(c.mesh, c.patch_u_bc)
        oc.set(P.F_U, c.U_i, c.U_b)
        oc.set(P.F_T, c.T_i, c.T_b)
        oc.set(P.F_ALPHA1, c.alpha1_i, c.alpha1_b)
        oc.set(P.F_P, c.p_i, c.p_b)
        if c.rotslip is not None:
            oc.bc_rotational_slip(c.rotslip)
        ocs.append(oc)
    cells = sum(c.mesh.nCells for c in cs)
    for _ in range(warmup):
        O.update_multi(ocs, cs[0].params, form=O.FAITHFUL)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.update_multi(ocs, cs[0].params, form=O.FAITHFUL)
    dt = time.perf_counter() - t0
    sample = (f"{steps} updates of a {nx}x{ny}x{lpt * threads} hex box ({cells} cells) split into {threads} z-slabs, "
              f"one per host thread; oracle faithful field-at-a-time restatement (OpenFOAM unavailable offline; its model "
              f"expressions are bit-identical to the reference's own sources, tests/test_reference_sources.py), "
              f"{workload} models")
    return cells * steps / dt / 1e6, threads, sample, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 2))
    # bound the arm to a few minutes: the full cfg2 box costs ~1 s per update on 16 cores
    steps = min(steps, 20)
    v, threads, sample, ms = cpu_arm(args.workload, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][1], "l2": "inputs larger than L2"},
        "cpu_baseline": {"value": round(v, 3), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- GPU arm -------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    from solidstateamfoam_b200 import capi
    from solidstateamfoam_b200 import params as P

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE")
    n_ranks = world
    torch.cuda.set_device(local)
    if n_ranks > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    t0 = time.time()
    case = build_case(args.workload, rank, n_ranks)
    mesh = case.mesh
    log(f"[rank {rank}] case {case.name}: {mesh.nCells} cells, {mesh.nInternalFaces} internal faces, "
        f"{mesh.nBoundaryFaces} boundary faces, built in {time.time() - t0:.1f}s")

    ctx = capi.Context(local)
    # the library launches on this torch stream, so torch.cuda.Event timings see its kernels
    # (a NULL handle would select the context's own stream, hence no legacy default stream here)
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    if n_ranks > 1:
        uid = [capi.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, n_ranks)
    ctx.load_case(case)
    ctx.set_cell_kernel(args.cell_kernel)
    flags = capi.FUSED_HALO_INPUTS if n_ranks > 1 else 0
    ctx.set_blocking(False)

    def barrier():
        if n_ranks > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ------------------------------------------------------------------
    for _ in range(args.warmup):
        ctx.update_fused(case.params, flags)
    barrier()
    ctx.profile_get()
    ctx.profile_enable(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        ctx.update_fused(case.params, flags)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    cells_ms, cells_n = ctx.profile_get()
    ctx.profile_enable(False)
    # the timed region lasts a few tens of ms, too short for nvidia-smi's sampling period: keep the very
    # same step running for ~1 s more so the clock record describes the GPU under this load.  The number
    # of extra steps must be identical on every rank (each step is a collective exchange when N > 1).
    tmax = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if n_ranks > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    n_extra = int(min(5000, max(20, 1000.0 / max(float(tmax.item()) / args.steps, 1e-3))))
    for i in range(n_extra):
        ctx.update_fused(case.params, flags)
        if i % 50 == 49:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed region + 1 s of the same step repeated"
    barrier()

    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    ncell = torch.tensor([float(mesh.nCells)], dtype=torch.float64, device="cuda")
    if n_ranks > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ncell, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    total_cells = float(ncell.item())
    value = total_cells * args.steps / (ms_max * 1e-3) / 1e6

    # ---- e2e: host buffers through ssam_update_fused_host ------------------------------------------
    def pinned(a):
        tt = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()
        return tt

    gi, gb = ctx.download(P.F_U)  # U with the tool-patch BC applied, as the solver would hold it
    ref_out_i, _ = ctx.download(case.outputs[0])
    layout_dev = "tile-compact" if ctx.layout_is_compact() else "caller order"
    if n_ranks == 1 and ctx.layout_is_compact() and os.environ.get("SSAM_E2E_SAME_CONTEXT") is None:
        # host-staged mode: PCIe dominates and the chunk pipeline needs contiguous cell ranges, so a caller that
        # keeps its fields on the host selects the caller-order layout (ssam_set_cell_layout) -- second context
        ctx_h = capi.Context(local)
        ctx_h.set_stream(stream.cuda_stream)
        ctx_h.set_cell_layout(P.LAYOUT_CALLER)
        ctx_h.mesh_set(case.mesh, case.patch_u_bc)
        ctx_h.set_blocking(False)
    else:
        ctx_h = ctx
    ins = {"U_i": pinned(gi), "U_b": pinned(gb), "T_i": pinned(case.T_i), "T_b": pinned(case.T_b),
           "alpha1_i": pinned(case.alpha1_i), "alpha1_b": pinned(case.alpha1_b), "p_i": pinned(case.p_i),
           "p_b": pinned(case.p_b)}
    outs = list(case.outputs)
    nOut = len(outs)
    out_i = [torch.empty(mesh.nCells, dtype=torch.float64).pin_memory() for _ in outs]
    out_b = [torch.empty(max(mesh.nBoundaryFaces, 1), dtype=torch.float64).pin_memory() for _ in outs]
    dp = C.POINTER(C.c_double)
    io = P.HostIO()
    for k, v in ins.items():
        setattr(io, k, C.cast(v.data_ptr(), dp))
    io.nOut = nOut
    of = (C.c_int32 * nOut)(*outs)
    oi = (dp * nOut)(*[C.cast(x.data_ptr(), dp) for x in out_i])
    ob = (dp * nOut)(*[C.cast(x.data_ptr(), dp) for x in out_b])
    io.outFields, io.out_i, io.out_b = of, oi, ob
    h2d = sum(v.numel() * 8 for v in ins.values())
    d2h = nOut * (mesh.nCells + mesh.nBoundaryFaces) * 8
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        ctx_h.update_fused_host(case.params, io)
    barrier()
    e0.record(stream)
    for _ in range(e2e_steps):
        ctx_h.update_fused_host(case.params, io)
    e1.record(stream)
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if n_ranks > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = total_cells * e2e_steps / (float(t2.item()) * 1e-3) / 1e6
    # the e2e outputs must be the update's results, not stale buffers
    assert np.array_equal(out_i[0].numpy(), ref_out_i, equal_nan=True), "e2e output mismatch"

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # algorithmic bytes of the cell kernel: everything except the friction-patch part (k_qfric)
        n_fric = int(mesh.patchSize[case.params.stick.fricPatch]) if case.params.stick.fricPatch >= 0 else 0
        bytes_cells = case.bytes_per_update - 40 * n_fric
        cells_avg_ms = cells_ms / max(cells_n, 1)
        achieved = bytes_cells / (cells_avg_ms * 1e-3) / 1e9
        traffic = None  # DRAM bytes per launch from the committed ncu capture of this workload and layout, if any
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = tr.get(args.workload, {}).get(layout_dev, {}).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": n_ranks, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload][1], "cells_per_gpu": mesh.nCells,
                       "parallelism": f"slab{n_ranks}" if n_ranks > 1 else "single",
                       "l2": "inputs larger than L2 (%.2f GB touched per update vs 126 MB L2)" %
                             (case.bytes_per_update / 1e9),
                       "cell_layout": layout_dev, "halo_transport": ctx.halo_transport(),
                       "e2e_cell_layout": "caller order (host-staged mode)" if ctx_h is not ctx else layout_dev},
            "roofline": {"bound": "hbm", "kernel": "k_cells_tiled" if args.cell_kernel == 1 else "k_cells", "achieved": round(achieved, 1), "peak": peak,
                         "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic if n_ranks == 1 else None,
                         "algorithmic_bytes_per_launch": bytes_cells, "kernel_ms": round(cells_avg_ms, 4),
                         "kernel_share_of_step": round(cells_avg_ms / (ms_total / args.steps), 4),
                         "peak_source": peak_src},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if n_ranks == 1 and not args.no_cpu_baseline:
            v, threads, sample, ms = cpu_arm(args.workload, 3, 1)
            line["cpu_baseline"] = {"value": round(v, 3), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": sample}
        print(json.dumps(line), flush=True)
    ctx.close()
    if n_ranks > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cell-kernel", type=int, default=1, choices=[0, 1],
                    help="1 = tiled, bulk-async staged (default); 0 = one thread per cell, direct loads")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.gpus > 1 and "RANK" not in os.environ:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                             "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
        run_gpu(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- Mcell-updates/s of the fused grad(U) -> nu -> mixture -> heat-source update.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5|cfg2|cfg4|cfg3] [--decomp block|slab]
                    [--impl reference]

Contract (one JSON line on stdout from rank 0):
  * a "step" = one fused update (ssam_update_fused) over the whole mesh of the workload;
  * default workload = BASELINE.json configs[4], the largest configuration that fits one GPU: the 64M-cell AFSD box
    (blockMesh hex 512x512x256, SheppardWright, constantRotationalSlip tool + constantSlipStick).  --gpus N is STRONG
    scaling of that same box: N structured blocks (1x2x1, 2x2x1, 2x2x2: the minimum-cut partition, up to three
    neighbours per rank) or N z-slabs (--decomp slab), one rank per GPU, processor-patch halos of U, temp, alpha1
    (inputs) and epsilonDotEq (outputs) exchanged every step;
  * value = cells of all ranks * K / max-over-ranks device time (CUDA events on the launching stream, barrier +
    synchronize on both sides), inputs resident in HBM;
  * parity = before anything is timed, every rank runs a small decomposed case (slab and ragged random partitions)
    through the same entry points and compares its partition with the multi-rank CPU oracle: integer halo maps and
    grad U / strain rate bit-exact, every other field <= 1e-11; a mismatch fails the run;
  * roofline = the dominant kernel (k_cells_tiled) timed live with its own CUDA events inside the library
    (ssam_profile_*), algorithmic bytes = 40*N_if + 52*N_bf + 128*N_c;
  * e2e = the same update through ssam_update_fused_host with pinned HOST buffers on the SAME context: H2D of U,
    temp, alpha1, p and D2H of the fields the solver consumes each step inside the timed region;
  * extra_workloads (N=1) = the other BASELINE configs (cfg2 FKS, cfg2 with general exponents, cfg3 polyhedral, cfg4
    NortonHoff) with kernel time and roofline fraction each;
  * cpu_baseline / --impl reference = the oracle's faithful field-at-a-time restatement of the reference (OpenFOAM
    itself is unavailable offline), one z-slab partition per host thread.
"""
from __future__ import annotations

import os as _os

# 1-5 MB halo messages: 4 NCCL channels are enough and leave more SMs to the interior tiles (profiles/r1_scaling.md)
_os.environ.setdefault("NCCL_MAX_NCHANNELS", "4")
import argparse
import ctypes as C
import faulthandler
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
    # name: (global dims, description, scaling with --gpus)
    "cfg5": ((512, 512, 256), "64M-cell AFSD box: blockMesh hex 512x512x256 (67.1M cells), SheppardWright, "
                              "constantRotationalSlip tool + constantSlipStick; N GPUs = N partitions of this box", "strong"),
    "cfg2": ((256, 256, 128), "blockMesh hex 256x256x128 (8.4M cells), FKS + cubic k(T), constantSlipStick", "weak"),
    "cfg2gen": ((256, 256, 128), "cfg2 with general FKS exponents (b2, c2 != 1: no pow(x,1) shortcuts)", "weak"),
    "cfg4": ((256, 256, 256), "hex 256^3 fluid region (16.8M cells), NortonHoff, qfric skipped", "single"),
    "cfg3": ((128, 128, 128), "polyhedral 5.7M cells (128^3 base, 2 levels of 2:1 refinement), SheppardWright, "
                              "exponentialRotationalSlip + exponentialSlipStick", "single"),
    "tiny": ((64, 64, 32), "debug: the cfg5 set-up on a 64x64x32 box", "strong"),
}
# what the solver reads back every step: TEqn.H:3-15,25 (epsilonDotEq, muEff, qvisc, cp, k) and UEqn's thermo.nu()
SOLVER_CONSUMED = ("EPSDOT", "MUEFF", "QVISC", "CP", "K", "NU")


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def build_case(workload: str, rank: int, n_ranks: int, decomp: str = "block", dims=None):
    from solidstateamfoam_b200 import cases

    dims = dims or WORKLOADS[workload][0]
    if workload in ("cfg5", "tiny"):
        return cases.cfg5_part(rank, n_ranks, decomp, dims=dims)
    if workload in ("cfg2", "cfg2gen"):  # N > 1: weak scaling, one 256x256x128 z-slab per rank (round-1 mode)
        nx, ny, nz = dims
        cell = 0.25e-3
        side = [-1] * 6
        if rank > 0:
            side[4] = rank - 1
        if rank < n_ranks - 1:
            side[5] = rank + 1
        box = ((0.0, 0.0, 0.0), (nx * cell, ny * cell, n_ranks * nz * cell))
        return cases.cfg2(dims=dims, side=tuple(side), rank=rank, origin=(0.0, 0.0, rank * nz * cell),
                          z0=0.6 * n_ranks * nz * cell, box=box,
                          visc1=cases.fks_general_exponents() if workload == "cfg2gen" else None)
    if n_ranks != 1:
        raise SystemExit(f"workload {workload} is single-GPU only")
    if workload == "cfg4":
        return cases.cfg4(dims=dims)
    if workload == "cfg3":
        return cases.cfg3(base=dims)
    raise SystemExit(f"unknown workload {workload}")


def config_of(args, n_ranks: int) -> dict:
    """identical for the GPU arm and the reference arm"""
    from solidstateamfoam_b200 import cases

    dims = WORKLOADS[args.workload][0]
    par = "single"
    if n_ranks > 1:
        if args.workload in ("cfg5", "tiny"):
            g = cases.block_grid(n_ranks, dims) if args.decomp == "block" else (1, 1, n_ranks)
            par = f"{args.decomp} {g[0]}x{g[1]}x{g[2]}"
        else:
            par = f"slab 1x1x{n_ranks} (one {dims[0]}x{dims[1]}x{dims[2]} slab per rank)"
    return {"workload": WORKLOADS[args.workload][1], "parallelism": par,
            "l2": "inputs larger than L2: ~249 B touched per cell-update vs 126 MB L2"}


# ---- host placement ------------------------------------------------------------------------------------
def pin_to_gpu_numa(local: int):
    """Bind this rank to the CPUs of its GPU's NUMA node before any pinned buffer is allocated, so the host staging
    buffers of the e2e leg are local to the PCIe root of the GPU (8 ranks otherwise share one socket's links)."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return None
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as fh:
            cpulist = fh.read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus ({cpulist})"
    except Exception:  # placement is an optimisation, never an error
        return None
    return None


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
def cpu_arm(workload: str, steps: int, warmup: int, threads: int | None = None, target_cells: float = 8.4e6):
    """The oracle's faithful field-at-a-time update (pass structure of the reference), one z-slab
    partition per host thread with in-memory halo copies (one MPI rank per core), on a bounded z-slab sample of the
    workload (~target_cells cells).  Returns (Mcell/s, threads, sample description, ms_per_step)."""
    from oracle import pyoracle as O
    from solidstateamfoam_b200 import cases
    from solidstateamfoam_b200 import params as P

    threads = threads or os.cpu_count() or 1
    nx, ny, nz = WORKLOADS[workload][0]
    if workload == "cfg3":
        nx = ny = 128
    lpt = max(1, min(int(round(target_cells / (nx * ny * threads))), max(nz // threads, 1)))
    if workload in ("cfg5", "tiny"):
        cs = [cases.cfg5_part(r, threads, "slab", dims=(nx, ny, lpt * threads)) for r in range(threads)]
    else:
        cs = [build_case("cfg2", r, threads, dims=(nx, ny, lpt)) for r in range(threads)]
    if workload in ("cfg4", "cfg3", "cfg2gen"):  # same slabs, that workload's constitutive models
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
    sample = (f"{steps} updates of a {nx}x{ny}x{lpt * threads} hex box ({cells} cells, a z-slab sample of the workload) "
              f"split into {threads} z-slabs, one per host thread; oracle faithful field-at-a-time restatement "
              f"(OpenFOAM unavailable offline; its model expressions are bit-identical to the reference's own sources, "
              f"tests/test_reference_sources.py), {workload} models")
    return cells * steps / dt / 1e6, threads, sample, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_ranks = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # every step is a bounded sample (~4M cells, ~0.25 s on 16 cores) so that K + W steps end within minutes
    v, threads, sample, ms = cpu_arm(args.workload, steps, warmup, target_cells=4.2e6 if steps + warmup > 12 else 8.4e6)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms, 3), "higher_is_better": True,
        "scaling": WORKLOADS[args.workload][2] if n_ranks > 1 else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args, n_ranks),
        "cpu_baseline": {"value": round(v, 3), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- parity record -------------------------------------------------------------------------------------
def parity_record(rank: int, n_ranks: int, local: int, dist, variant: int):
    """A small decomposed case through the same entry points, against the multi-rank CPU oracle (checker use of
    oracle/, like tests/multigpu_parity.py).  Raises on mismatch."""
    from oracle import pyoracle as O
    from solidstateamfoam_b200 import capi, cases
    from solidstateamfoam_b200 import params as P

    bit_exact = (P.F_GRADU, P.F_EPSDOT, P.F_RHO, P.F_CP, P.F_NU2, P.F_U, P.F_T, P.F_ALPHA1)
    worst, exact_names, decomps = 0.0, set(), []
    for decomp in (("slab", "random") if n_ranks > 1 else ("serial",)):
        parts = cases.parity_parts(n_ranks, decomp)
        ocs = []
        for c in parts:
            oc = O.OracleCase(c.mesh, c.patch_u_bc)
            oc.set(P.F_U, c.U_i, c.U_b)
            oc.set(P.F_T, c.T_i, c.T_b)
            oc.set(P.F_ALPHA1, c.alpha1_i, c.alpha1_b)
            oc.set(P.F_P, c.p_i, c.p_b)
            oc.bc_rotational_slip(c.rotslip)
            ocs.append(oc)
        if n_ranks > 1:
            O.halo_exchange(ocs, P.F_T)
            O.halo_exchange(ocs, P.F_ALPHA1)
        O.update_multi(ocs, parts[0].params, form=O.FAITHFUL)
        case, oc = parts[rank], ocs[rank]
        ctx = capi.Context(local)
        try:
            if n_ranks > 1:
                uid = [capi.Context.comm_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(uid, src=0)
                ctx.comm_init(uid[0], rank, n_ranks)
            ctx.load_case(case)
            ctx.set_cell_kernel(variant)
            for which in (P.ADDR_HALO_SEND_CELLS, P.ADDR_HALO_PATCH_OFFSETS, P.ADDR_EXTRA_FACE, P.ADDR_EXTRA_OTHER):
                if ctx.addressing(which).tobytes() != oc.addressing(which).tobytes():
                    raise AssertionError(f"parity: integer table {P.ADDR_NAMES[which]} differs on rank {rank} ({decomp})")
            flags = capi.FUSED_WRITE_GRAD | (capi.FUSED_HALO_INPUTS if n_ranks > 1 else 0)
            for _ in range(2):  # twice: the second update reuses every cached / double-buffered piece of the exchange
                ctx.update_fused(case.params, flags)
            for f in case.outputs + [P.F_GRADU, P.F_U, P.F_T, P.F_ALPHA1]:
                gi, gb = ctx.download(f)
                got, ref = np.concatenate([gi, gb]), oc.field(f)
                if f in bit_exact:
                    same = (got == ref) | (np.isnan(got) & np.isnan(ref))
                    if not same.all():
                        raise AssertionError(f"parity: {P.FIELD_NAMES[f]} not bit-exact on rank {rank} ({decomp}): "
                                             f"{np.count_nonzero(~same)} of {same.size} values differ")
                    exact_names.add(P.FIELD_NAMES[f])
                else:
                    d = np.abs(got - ref)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        r = np.where(d == 0, 0.0, d / np.abs(ref))
                    r = np.where(np.isnan(got) & np.isnan(ref), 0.0, r)
                    w = float(np.nan_to_num(r, nan=np.inf).max()) if r.size else 0.0
                    if not w <= 1e-11:
                        raise AssertionError(f"parity: {P.FIELD_NAMES[f]} rel err {w:.3e} > 1e-11 on rank {rank} ({decomp})")
                    worst = max(worst, w)
            transport = ctx.halo_transport()
        finally:
            ctx.close()
        decomps.append(decomp)
    if n_ranks > 1:
        import torch

        t = torch.tensor([worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        worst = float(t.item())
    return {"max_rel": worst, "tolerance": 1e-11, "bit_exact_fields": sorted(exact_names), "ranks": n_ranks,
            "decomps": decomps, "cells_per_rank": "432-1920", "oracle": "multi-rank CPU restatement (oracle/), faithful form",
            "halo_transport": transport}


# ---- GPU arm -------------------------------------------------------------------------------------------
def time_updates(ctx, case, flags, steps, warmup, stream, barrier, torch):
    """(total ms over `steps` updates, launches, kernel ms sum, kernel launches)"""
    for _ in range(warmup):
        ctx.update_fused(case.params, flags)
    barrier()
    ctx.profile_get()
    ctx.profile_enable(True)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(steps):
        ctx.update_fused(case.params, flags)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    cells_ms, cells_n = ctx.profile_get()
    ctx.profile_enable(False)
    return ms_total, launches, cells_ms, cells_n


def cell_kernel_bytes(case):
    """algorithmic bytes of the cell kernel: everything except the friction-patch part (k_qfric)"""
    n_fric = int(case.mesh.patchSize[case.params.stick.fricPatch]) if case.params.stick.fricPatch >= 0 else 0
    return case.bytes_per_update - 40 * n_fric


def ncu_record(workload: str, layout: str) -> dict:
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return tr.get(workload, {}).get(layout, {}) or {}
    except (OSError, ValueError):
        return {}


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
    numa = pin_to_gpu_numa(local)
    torch.cuda.set_device(local)
    if n_ranks > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if n_ranks > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def phase(name, limit_s):
        """one stderr line per phase, and a watchdog: a phase that exceeds its limit (a stalled collective) dumps every
        thread's Python stack and ends the process instead of hanging the caller"""
        if rank == 0 or os.environ.get("SSAM_BENCH_VERBOSE"):
            log(f"[rank {rank}] phase: {name}")
        try:
            faulthandler.cancel_dump_traceback_later()
            if limit_s:
                faulthandler.dump_traceback_later(limit_s, exit=True, file=sys.__stderr__)
        except (ValueError, OSError, AttributeError):  # stderr without a file descriptor: no watchdog, never an error
            pass

    # ---- parity first: nothing is timed unless the decomposed results match the oracle -------------------
    parity = None
    phase("parity record", 600)
    if not args.no_parity:
        t0 = time.time()
        parity = parity_record(rank, n_ranks, local, dist, args.cell_kernel)
        log(f"[rank {rank}] parity ok ({parity['decomps']}, max rel {parity['max_rel']:.2e}) in {time.time() - t0:.1f}s")

    phase("mesh generation + upload", 900)
    t0 = time.time()
    case = build_case(args.workload, rank, n_ranks, args.decomp)
    mesh = case.mesh
    n_proc = int(sum(mesh.patchSize[p] for p in range(mesh.nPatches) if mesh.patchKind[p] == 1))
    log(f"[rank {rank}] case {case.name}: {mesh.nCells} cells, {mesh.nInternalFaces} internal faces, "
        f"{mesh.nBoundaryFaces} boundary faces ({n_proc} on processor patches), built in {time.time() - t0:.1f}s")

    t0 = time.time()
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
    log(f"[rank {rank}] mesh + fields on the device in {time.time() - t0:.1f}s")

    # ---- device-resident timing ------------------------------------------------------------------
    sampler = ClockSampler(local)
    phase("warm-up updates", 300)
    for _ in range(args.warmup):
        ctx.update_fused(case.params, flags)
    barrier()
    phase("timed updates + sustained tail", 300)
    if rank == 0:
        sampler.start()
    ms_total, launches, cells_ms, cells_n = time_updates(ctx, case, flags, args.steps, 0, stream, barrier, torch)
    # The timed region can be shorter than nvidia-smi's sampling period: keep the very same step running for ~1 s
    # more so the clock record describes the GPU under this load, and time that tail too ("sustained").  The number
    # of extra steps must be identical on every rank (each step is a collective exchange when N > 1).
    tmax = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if n_ranks > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    n_extra = int(min(5000, max(20, 1000.0 / max(float(tmax.item()) / args.steps, 1e-3))))
    es0, es1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    es0.record(stream)
    for i in range(n_extra):
        ctx.update_fused(case.params, flags)
        if i % 50 == 49:
            torch.cuda.synchronize()
    es1.record(stream)
    barrier()
    ms_sustained = es0.elapsed_time(es1)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "timed region + the ~1 s sustained tail of the same step"

    t = torch.tensor([ms_total, ms_sustained], dtype=torch.float64, device="cuda")
    ncell = torch.tensor([float(mesh.nCells)], dtype=torch.float64, device="cuda")
    kms = torch.tensor([cells_ms / max(cells_n, 1)], dtype=torch.float64, device="cuda")
    if n_ranks > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ncell, op=dist.ReduceOp.SUM)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_max, ms_sus = float(t[0].item()), float(t[1].item())
    total_cells = float(ncell.item())
    value = total_cells * args.steps / (ms_max * 1e-3) / 1e6
    sustained = total_cells * n_extra / (ms_sus * 1e-3) / 1e6

    layout_dev = "tile-compact" if ctx.layout_is_compact() else "caller order"

    # ---- e2e: host buffers through ssam_update_fused_host, same context / same cell layout ---------------
    def e2e_leg():
        def pinned(a):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).pin_memory()

        gi, gb = ctx.download(P.F_U)  # U with the tool-patch BC applied, as the solver would hold it
        ins = {"U_i": pinned(gi), "U_b": pinned(gb), "T_i": pinned(case.T_i), "T_b": pinned(case.T_b),
               "alpha1_i": pinned(case.alpha1_i), "alpha1_b": pinned(case.alpha1_b), "p_i": pinned(case.p_i),
               "p_b": pinned(case.p_b)}
        del gi, gb
        outs = [getattr(P, "F_" + n) for n in SOLVER_CONSUMED]
        ref_out_i, _ = ctx.download(outs[0])
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
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            ctx.update_fused_host(case.params, io)
        barrier()
        e0.record(stream)
        for _ in range(e2e_steps):
            ctx.update_fused_host(case.params, io)
        e1.record(stream)
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if n_ranks > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_value = total_cells * e2e_steps / (float(t2.item()) * 1e-3) / 1e6
        # the e2e outputs must be the update's results, not stale buffers
        assert np.array_equal(out_i[0].numpy(), ref_out_i, equal_nan=True), "e2e output mismatch"
        del ref_out_i
        return {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "outputs": list(SOLVER_CONSUMED), "cell_layout": layout_dev,
                "note": "same context and cell layout as `value`; D2H = the fields the solver consumes per step "
                        "(TEqn.H:3-15,25 and thermo.nu())"}

    phase("fp64 ceiling", 300)
    fp64_peak = ctx.fp64_peak() if hasattr(ctx, "fp64_peak") else None
    transport = ctx.halo_transport()
    nvl = ctx.halo_bytes_per_update() if hasattr(ctx, "halo_bytes_per_update") else None

    # ---- the line as far as the device-resident measurement goes (rank 0); the later legs add to it ----------------
    line, peak = None, None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        bytes_cells = cell_kernel_bytes(case)
        cells_avg_ms = float(kms.item())
        achieved = bytes_cells / (cells_avg_ms * 1e-3) / 1e9
        kname = {0: "k_cells", 1: "k_cells_tiled", 2: "k_cells_pipe"}[args.cell_kernel]
        rec = ncu_record(args.workload, layout_dev) if n_ranks == 1 else {}
        roof = {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": rec.get("dram_bytes_per_launch"),
                "algorithmic_bytes_per_launch": bytes_cells, "kernel_ms": round(cells_avg_ms, 4),
                "kernel_share_of_step": round(cells_avg_ms / (ms_max / args.steps), 4), "peak_source": peak_src}
        if fp64_peak:
            roof["fp64_peak_ginstr_s"] = round(fp64_peak / 1e9, 1)
            per_cell = rec.get("fp64_thread_instr_per_cell")
            if per_cell:
                roof["fp64_frac"] = round(per_cell * mesh.nCells / (cells_avg_ms * 1e-3) / fp64_peak, 4)
                roof["fp64_thread_instr_per_cell"] = per_cell
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": n_ranks, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_max / args.steps, 4), "higher_is_better": True,
            "scaling": WORKLOADS[args.workload][2] if n_ranks > 1 else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_of(args, n_ranks),
            "detail": {"cells_total": int(total_cells), "cells_rank0": mesh.nCells, "processor_faces_rank0": n_proc,
                       "cell_layout": layout_dev, "halo_transport": transport, "host_numa_binding": numa,
                       "nvlink_bytes_per_update_rank0": nvl},
            "sustained": {"value": round(sustained, 2), "unit": UNIT, "steps": n_extra,
                          "ms_per_step": round(ms_sus / n_extra, 4)},
            "parity": parity,
            "roofline": roof,
            "e2e": {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": "not run"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }

    class leg:
        """One of the legs after the timed region.  On one GPU a leg that exceeds its limit must not cost the line: a
        timer prints what has been measured so far (naming the leg) and ends the process.  With several ranks the
        phase watchdog applies instead (a rank cannot leave the collective sequence on its own)."""

        def __init__(self, name, limit_s):
            self.name, self.limit_s, self.timer = name, limit_s, None

        def expire(self):
            line["aborted_leg"] = f"{self.name}: no result within {self.limit_s} s"
            print(json.dumps(line), flush=True)
            os._exit(0)

        def __enter__(self):
            if n_ranks == 1:
                phase(self.name, 0)
                self.timer = threading.Timer(self.limit_s, self.expire)
                self.timer.daemon = True
                self.timer.start()
            else:
                phase(self.name, self.limit_s)
            return self

        def __exit__(self, *exc):
            if self.timer:
                self.timer.cancel()
            return False

    with leg("e2e leg (host buffers)", args.leg_limit or 600):
        try:
            e2e = e2e_leg()
        except Exception as exc:  # the device-resident line must still be printed; the failure is part of it
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                   "error": f"{type(exc).__name__}: {exc}"[:400]}
            log(f"[rank {rank}] e2e leg failed: {exc}")
            if n_ranks > 1:
                raise  # a rank that left the collective sequence cannot be resynchronised
        if line is not None:
            line["e2e"] = e2e
    phase("teardown", 300)
    barrier()  # no rank frees buffers a neighbour may still be using
    ctx.close()
    phase("report", 0 if n_ranks == 1 else 300)

    if rank == 0:
        if n_ranks == 1 and not args.no_extra:
            with leg("extra workloads", args.leg_limit or 900):
                try:
                    line["extra_workloads"] = extra_workloads(args, stream, torch, peak)
                except Exception as exc:
                    line["extra_workloads"] = {"error": f"{type(exc).__name__}: {exc}"[:400]}
        if n_ranks == 1 and not args.no_cpu_baseline:
            with leg("cpu baseline", args.leg_limit or 900):
                try:
                    v, threads, sample, ms = cpu_arm(args.workload, 3, 1)
                    line["cpu_baseline"] = {"value": round(v, 3), "unit": UNIT, "cores": threads, "kind": "port",
                                            "sample": sample}
                except Exception as exc:
                    line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port",
                                            "sample": f"failed: {type(exc).__name__}: {exc}"[:400]}
        print(json.dumps(line), flush=True)
    phase("done", 0)
    if n_ranks > 1:
        dist.barrier()
        dist.destroy_process_group()


def extra_workloads(args, stream, torch, peak):
    """the other BASELINE configs on one GPU: kernel time, roofline fraction, recorded DRAM traffic"""
    from solidstateamfoam_b200 import capi

    out = []
    for wl in ("cfg2", "cfg2gen", "cfg3", "cfg4"):
        if wl == args.workload:
            continue
        t0 = time.time()
        case = build_case(wl, 0, 1)
        ctx = capi.Context(torch.cuda.current_device())
        ctx.set_stream(stream.cuda_stream)
        ctx.load_case(case)
        ctx.set_cell_kernel(args.cell_kernel)
        ctx.set_blocking(False)
        steps = 20
        ms_total, launches, cells_ms, cells_n = time_updates(ctx, case, 0, steps, 4, stream, torch.cuda.synchronize, torch)
        layout = "tile-compact" if ctx.layout_is_compact() else "caller order"
        ctx.close()
        k = cells_ms / max(cells_n, 1)
        nbytes = cell_kernel_bytes(case)
        rec = ncu_record(wl, layout)
        out.append({"workload": WORKLOADS[wl][1], "name": wl, "cells": case.mesh.nCells,
                    "value": round(case.mesh.nCells * steps / (ms_total * 1e-3) / 1e6, 2), "unit": UNIT,
                    "ms_per_step": round(ms_total / steps, 4), "kernel_ms": round(k, 4),
                    "frac": round(nbytes / (k * 1e-3) / 1e9 / peak, 4), "algorithmic_bytes_per_launch": nbytes,
                    "traffic": rec.get("dram_bytes_per_launch"), "cell_layout": layout, "steps": steps})
        log(f"[extra] {wl}: {out[-1]['value']} {UNIT}, kernel {k:.4f} ms, frac {out[-1]['frac']}, "
            f"{time.time() - t0:.1f}s incl. set-up")
        del case
    return out


def make_parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg5", choices=list(WORKLOADS))
    ap.add_argument("--decomp", default="block", choices=["block", "slab"],
                    help="partition of the cfg5 box at N > 1: structured minimum-cut blocks (default) or z-slabs")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra_workloads list (N=1)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity record (experiments only)")
    ap.add_argument("--cell-kernel", type=int, default=1, choices=[0, 1, 2],
                    help="1 = tiled, one CTA per tile (default); 2 = persistent self-pipelined tiles; 0 = one thread per "
                         "cell, direct loads")
    ap.add_argument("--leg-limit", type=float, default=0.0,
                    help="seconds allowed to each leg after the timed region (default: 600 e2e, 900 extras / CPU arm)")
    return ap


def main():
    args = make_parser().parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        if args.gpus > 1 and "RANK" not in os.environ:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                             "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
        run_gpu(args)


if __name__ == "__main__":
    main()

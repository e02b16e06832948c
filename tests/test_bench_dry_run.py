"""bench.py's GPU arm executed WITHOUT a GPU, against stand-ins for the device context and the torch.cuda calls: the
control flow, the JSON contract (every key the driver reads) and the safety nets (a failing or stalling leg after the
timed region must not cost the line) are host logic and are checked here; the numbers are meaningless by construction.
The real arm is covered on the GPU box by the driver's own bench run."""
import ctypes as C
import importlib.util
import json
import os
import sys
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeEvent:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return max((other.t - self.t) * 1e3, 1e-3)


class FakeStream:
    cuda_stream = 0x1234

    def __init__(self, device=None):
        pass


class FakeContext:
    """what bench.py calls on capi.Context; update_fused_host honours the HostIO contract (writes every output)"""
    stall_host = False
    fail_host = False

    def __init__(self, device):
        self.n_launch = 0
        self.case = None
        self.prof = [0.0, 0]
        self.profiling = False

    @staticmethod
    def comm_unique_id():
        return b"\0" * 128

    def comm_init(self, uid, rank, n_ranks):
        assert len(uid) == 128 and 0 <= rank < n_ranks

    def set_stream(self, s):
        assert s != 0

    def load_case(self, case):
        self.case = case

    def set_cell_kernel(self, v):
        pass

    def set_blocking(self, b):
        pass

    def update_fused(self, params, flags):
        self.n_launch += 3
        if self.profiling:
            self.prof[0] += 0.5
            self.prof[1] += 1

    def profile_get(self):
        r = tuple(self.prof)
        self.prof = [0.0, 0]
        return r

    def profile_enable(self, on):
        self.profiling = bool(on)

    def launch_count(self):
        return self.n_launch

    def layout_is_compact(self):
        return True

    def download(self, field):
        m = self.case.mesh
        from solidstateamfoam_b200 import params as P

        comps = 3 if field == P.F_U else 1
        return np.full(m.nCells * comps, 1.5), np.full(max(m.nBoundaryFaces, 1) * comps, 2.5)

    def update_fused_host(self, params, io):
        if FakeContext.fail_host:
            raise RuntimeError("SSAM_ERR_CUDA: injected failure")
        if FakeContext.stall_host:
            time.sleep(3600)
        m = self.case.mesh
        src = np.full(m.nCells, 1.5)
        for k in range(io.nOut):
            C.memmove(io.out_i[k], src.ctypes.data, m.nCells * 8)

    def fp64_peak(self):
        return 17.4e12

    def halo_transport(self):
        return "none"

    def halo_bytes_per_update(self):
        return 0

    def close(self):
        pass


def load_bench(monkeypatch):
    import torch

    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from solidstateamfoam_b200 import capi

    monkeypatch.setattr(capi, "Context", FakeContext)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "set_stream", lambda s: None)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    real_tensor = torch.tensor
    monkeypatch.setattr(torch, "tensor", lambda *a, **k: real_tensor(*a, **{x: y for x, y in k.items() if x != "device"}))
    for v in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(v, raising=False)
    monkeypatch.setattr(bench, "pin_to_gpu_numa", lambda local: None)
    return bench


REQUIRED = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks")


def run(bench, capsys, *argv):
    args = bench.make_parser().parse_args(list(argv))
    bench.run_gpu(args)
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(out) == 1, out
    return json.loads(out[0])


def test_gpu_arm_control_flow_and_json_contract(monkeypatch, capsys):
    bench = load_bench(monkeypatch)
    FakeContext.fail_host = FakeContext.stall_host = False
    line = run(bench, capsys, "--workload", "tiny", "--steps", "4", "--warmup", "3", "--no-parity", "--no-extra",
               "--no-cpu-baseline")
    for k in REQUIRED:
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == 4 and line["warmup"] == 3
    assert line["higher_is_better"] is True and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["config"]["workload"].startswith("debug")
    r = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert r["kernel_ms"] == 0.5  # the stand-in's per-update kernel time: the average over the timed steps only
    e = line["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["outputs"] == list(bench.SOLVER_CONSUMED)
    assert line["gpu_launches"] == 3 * 4
    assert line["sustained"]["steps"] >= 20


def test_failing_e2e_leg_still_prints_the_line(monkeypatch, capsys):
    bench = load_bench(monkeypatch)
    FakeContext.fail_host, FakeContext.stall_host = True, False
    try:
        line = run(bench, capsys, "--workload", "tiny", "--steps", "2", "--warmup", "3", "--no-parity", "--no-extra",
                   "--no-cpu-baseline")
    finally:
        FakeContext.fail_host = False
    assert line["value"] > 0
    assert line["e2e"]["value"] is None and "injected failure" in line["e2e"]["error"]


def test_stalled_leg_prints_the_line_and_ends_the_process(tmp_path):
    """a leg that never returns: the timer prints the device-resident line (naming the leg) and exits 0"""
    script = tmp_path / "stall.py"
    script.write_text(f"""
import sys, types
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
import pytest
import test_bench_dry_run as T
mp = pytest.MonkeyPatch()
bench = T.load_bench(mp)
T.FakeContext.stall_host = True
args = bench.make_parser().parse_args(["--workload", "tiny", "--steps", "2", "--warmup", "3", "--no-parity",
                                       "--no-extra", "--no-cpu-baseline", "--leg-limit", "2"])
bench.run_gpu(args)
""")
    import subprocess

    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    out = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["value"] > 0 and "e2e leg" in line["aborted_leg"] and line["e2e"]["value"] is None


def test_reference_arm_honours_steps_and_warmup(monkeypatch, capsys):
    bench = load_bench(monkeypatch)
    args = bench.make_parser().parse_args(["--impl", "reference", "--workload", "tiny", "--steps", "2", "--warmup", "1"])
    bench.run_reference(args)
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    line = json.loads(out[-1])
    assert line["impl"] == "reference" and line["steps"] == 2 and line["warmup"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"] == bench.config_of(args, 1)


def test_two_rank_flow_over_gloo(tmp_path):
    """the N > 1 branch (rendezvous, unique-id broadcast, block partition of the box, max-over-ranks reductions, one
    line from rank 0 only) with two real processes; torch.distributed runs over gloo, the device stays a stand-in"""
    import subprocess

    script = tmp_path / "two_ranks.py"
    script.write_text(f"""
import sys
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
import os
import pytest
import torch.distributed as dist
world, rank, local = os.environ["WORLD_SIZE"], os.environ["RANK"], os.environ["LOCAL_RANK"]
import test_bench_dry_run as T
mp = pytest.MonkeyPatch()
bench = T.load_bench(mp)
os.environ.update(WORLD_SIZE=world, RANK=rank, LOCAL_RANK=local)
real_init = dist.init_process_group
mp.setattr(dist, "init_process_group", lambda backend, **kw: real_init("gloo"))
args = bench.make_parser().parse_args(["--gpus", "2", "--workload", "tiny", "--steps", "3", "--warmup", "3",
                                       "--no-parity"])
bench.run_gpu(args)
""")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    out = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(out) == 1, out
    line = json.loads(out[0])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong"
    assert line["config"]["parallelism"] == "block 1x2x1"
    assert line["detail"]["cells_total"] == 64 * 64 * 32 and line["detail"]["cells_rank0"] == 64 * 32 * 32
    assert line["detail"]["processor_faces_rank0"] == 64 * 32
    assert line["e2e"]["value"] > 0
    assert "extra_workloads" not in line and "cpu_baseline" not in line

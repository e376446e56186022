"""Host-side pieces of bench.py that run without a GPU: the contract's names, the grid shapes, the clock sampler's summary when NVML is
absent, and the refusal to run without a B200 (there is no CPU fallback behind the benchmark)."""
import importlib.util
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_workload_name_and_grid_shapes():
    b = _bench()
    cfg = b.workload_config(10000, 1)
    assert "dgemm m=n=k=10000" in cfg["workload"] and "configs[1]" in cfg["workload"] and "2 streams" in cfg["workload"]
    assert [b.grid_shape(n) for n in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
    assert b.host_cores() >= 1
    # N > 1: the square problem weak-scaled by work: 2e12 flop per GPU within 0.1 %
    for world in (1, 2, 4, 8):
        S = b.global_size(world, 10000)
        assert abs(2.0 * S**3 / world - 2e12) <= 2e9, (world, S)
        assert f"m=n=k={S} " in b.workload_config(S, world)["workload"]
        pr, pc = b.grid_shape(world)
        rows = [b.share(S, pr, g) for g in range(pr)]
        assert rows[0][0] == 0 and rows[-1][1] == S and all(rows[g][1] == rows[g + 1][0] for g in range(pr - 1))


def test_clock_sampler_summary_without_nvml():
    b = _bench()
    cs = b.ClockSampler(0)
    cs.nv = None  # as on a box without NVML
    with cs:
        pass
    s = cs.summary()
    assert s["reasons"] == [] and s["sm_mhz"] is None and s["samples"] == 0
    cs.samples, cs.reasons, cs.power = [0, 1965, 1965, 1350], {"SwPowerCap", "HwThermalSlowdown"}, [300.0, 640.5]
    s = cs.summary()
    assert s["sm_mhz"] == 1965 and s["reasons"] == ["hw_thermal_slowdown", "sw_power_cap"] and s["power_w_max"] == 640.5


def test_bench_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1", "--no-cpu-baseline"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


MOCK_RUNNER = r'''
import runpy, sys, types
import numpy as np

# ---- a stand-in for torch: just enough surface for bench.py's single-GPU flow ----
torch = types.ModuleType("torch")
torch.float64 = "f64"; torch.int64 = "i64"
class _Stream: cuda_stream = 0
class _Event:
    clock = [0.0]
    def __init__(self, enable_timing=False): self.t = 0.0
    def record(self, st=None): _Event.clock[0] += 1.0; self.t = _Event.clock[0]
    def elapsed_time(self, other): return max(other.t - self.t, 0.5)
class _Dev:
    def __init__(self, n): self.a = np.zeros(n)
    def copy_(self, x): return self
    def data_ptr(self): return 0
    def view(self, *s): return self
    def t(self): return self
cuda = types.SimpleNamespace(is_available=lambda: True, set_device=lambda i: None, current_stream=lambda: _Stream(), Event=_Event,
                             synchronize=lambda: None, empty_cache=lambda: None)
torch.cuda = cuda
torch.empty = lambda n, dtype=None, device=None: _Dev(n)
torch.from_numpy = lambda x: x
torch.matmul = lambda a, b: None
torch.device = lambda *a: None
sys.modules["torch"] = torch

# ---- a stand-in for the library: gemm really multiplies (numpy), so the benchmark's linearity check has something to check ----
tmm = types.ModuleType("tiled_mm_b200")
state = {"launches": 0}
class _Stats: h2d_bytes = 2 * 64 * 64 * 8; d2h_bytes = 64 * 64 * 8; peer_bytes = 0
class _Ctx:
    def last_stats(self): return _Stats()
    def close(self): pass
def gemm(ctx, ta, tb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, pin_host_buffers=False, copy_c_back=True):
    A = np.asarray(a).reshape(k, lda)[:, :m].T
    B = np.asarray(b).reshape(n, ldb)[:, :k].T
    np.asarray(c).reshape(n, ldc)[:, :m] = (alpha * (A @ B)).T
    state["launches"] += 7
def device_gemm(*a, **k): state["launches"] += 1
tmm.device_count = lambda: 1
tmm.malloc_pinned = lambda dtype, count: np.zeros(count, dtype)
tmm.make_context = lambda *a, **k: _Ctx()
tmm.gemm = gemm
tmm.device_gemm = device_gemm
tmm.total_kernel_launches = lambda: state["launches"]
tmm.probe_fp64_peak = lambda: 37.0
tmm.probe_host_links = lambda ids=None, n_devices=None, nbytes=0: [(50.0, 50.0) for _ in (ids or [0])]
sys.modules["tiled_mm_b200"] = tmm

sys.argv = ["bench.py", "--size", "64", "--steps", "2", "--warmup", "1", "--no-cpu-baseline"]
runpy.run_path(BENCH, run_name="__main__")
'''


def test_bench_main_flow_on_stand_ins(tmp_path):
    """The whole single-GPU flow of bench.py (both timed legs, the re-measure decision, the JSON line) executed on stand-ins for torch and
    for the library - guards the benchmark's own code against slips that would only show on the GPU box."""
    import json
    runner = tmp_path / "run_bench.py"
    runner.write_text("BENCH = %r\n" % str(ROOT / "bench.py") + MOCK_RUNNER)
    r = subprocess.run([sys.executable, str(runner)], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
                "gpu_launches", "roofline", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 3 and line["dtype"] == "f64" and line["gpu_launches"] == 14
    assert line["value"] == line["e2e"]["value"] and line["ms_per_step"] == line["e2e"]["ms_per_step"], "value must be the host-to-host measurement the metric names"
    assert line["roofline"]["peak"] == 37.0 and line["e2e"]["host_roofline"]["aggregate_h2d_gbs"] == 50.0
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert line["clocks"]["remeasured_after_slowdown"] is False and "workload" in line["config"]

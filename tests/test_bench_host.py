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
    name = b.workload_name(10000)
    assert "dgemm m=n=k=10000" in name and "configs[1]" in name and "2 streams" in name
    assert [b.grid_shape(n) for n in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]
    assert b.host_cores() >= 1


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

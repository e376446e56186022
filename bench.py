#!/usr/bin/env python
"""bench.py - headline benchmark of the hot path gpu::make_context -> gpu::gemm.

Workload (BASELINE.json configs[1], the README miniapp): dgemm m=n=k=10000, NN, alpha=1, beta=0, pinned host buffers,
tile hints 5000^3, 2 streams.  A "step" is one gemm.

  value   device-resident DGEMM throughput (operands already in HBM, one kernel launch per step, CUDA events)
  e2e     the same GEMM through the public call with HOST buffers (H2D of A,B and D2H of C inside the timed region) -
          this is the library's actual product and the headline against the reference arm
  roofline  FP64 tensor (DMMA) bound for the dominant kernel;  cpu_baseline  host BLAS dgemm on the box's cores

N > 1 (torchrun, one rank per GPU): weak scaling - rank (i,j) of a p_r x p_c grid owns one 10000 x 10000 block of C of the
global (p_r*10000) x (p_c*10000) x 10000 product; each rank uploads only its 1/p_c slice of the A row-panel and 1/p_r
slice of the B column-panel over its own PCIe link and the slices are all-gathered over NVLink (NCCL) before the local GEMM.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size S]
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

FP64_PEAK_TFLOPS = 36.9  # DMMA.8x8x4 issue microbenchmark on this pool's B200 (profiles/r1_probe_b200.txt); 148 SM x 64 FMA/clk x 1.965 GHz = 37.2
PCIE_H2D_GBS, PCIE_D2H_GBS = 55.6, 57.0  # pinned cudaMemcpyAsync 1 GiB, same probe


class ClockSampler:
    """Samples SM clock / throttle reasons during the timed region (pynvml; falls back to nvidia-smi)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.power = index, [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {}
        for nm in dir(nv):
            if nm.startswith("nvmlClocksEventReason") or nm.startswith("nvmlClocksThrottleReason"):
                v = getattr(nv, nm)
                if isinstance(v, int) and v:
                    names[v] = nm.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", "")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit and nm not in ("GpuIdle", "None", "All"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        busy = sorted(s for s in self.samples if s > 0)
        rename = {"SwPowerCap": "sw_power_cap", "HwSlowdown": "hw_slowdown", "HwThermalSlowdown": "hw_thermal_slowdown",
                  "SwThermalSlowdown": "sw_thermal_slowdown", "HwPowerBrakeSlowdown": "hw_power_brake", "ApplicationsClocksSetting": "app_clocks",
                  "SyncBoost": "sync_boost", "DisplayClockSetting": "display_clock"}
        return {"sm_mhz": (busy[len(busy) // 2] if busy else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(rename.get(r, r) for r in self.reasons), "power_w_max": (max(self.power) if self.power else None),
                "samples": len(self.samples)}


def fill_uniform(arr: np.ndarray, seed: int) -> None:
    """uniform(-1,1) doubles, generated in slabs (counter-based seeding so ranks/slabs are independent)."""
    slab = 1 << 22
    for i, off in enumerate(range(0, arr.size, slab)):
        rng = np.random.default_rng([seed, i])
        n = min(slab, arr.size - off)
        arr[off:off + n] = rng.random(n) * 2.0 - 1.0


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline_dgemm(size: int) -> dict:
    """Host BLAS dgemm (numpy -> OpenBLAS, all cores) on the workload itself (m = n = k = size), repeated until about 10 s of CPU
    work have been timed (at most 8 runs; one run if the box's cores need longer than that for a single product)."""
    m = n = size
    k = size
    rng = np.random.default_rng(0)
    a = np.asfortranarray(rng.random((m, k)) - 0.5)
    b = np.asfortranarray(rng.random((k, n)) - 0.5)
    np.dot(a[:256], b[:, :256])
    runs, total = 0, 0.0
    while total < 10.0 and runs < 8:  # about 10-20 s of CPU work
        t0 = time.perf_counter()
        c = np.dot(a, b)
        total += time.perf_counter() - t0
        runs += 1
        assert np.isfinite(c[0, 0])
    dt = total / runs
    return {"value": round(2.0 * m * n * k / dt * 1e-12, 4), "unit": "TFLOP/s", "cores": host_cores(), "kind": "port",
            "sample": f"host BLAS dgemm (numpy/OpenBLAS, all {host_cores()} hardware threads) on the full workload {m}x{n}x{k}, mean of {runs} runs, {total:.1f} s of CPU work"}


class NumaLocal:
    """While active, this process runs on the CPUs NVML reports as local to GPU `index`, so that the pinned host buffers
    allocated (and first touched) inside land on that GPU's NUMA node: with 8 ranks streaming over 8 PCIe links, buffers on
    the remote socket would halve the achievable host-link bandwidth.  The previous mask is restored on exit (pages stay put)."""

    def __init__(self, index: int):
        self.index, self.prev, self.bound = index, None, None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            local = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
            self.prev = os.sched_getaffinity(0)
            want = local & self.prev
            if want and want != self.prev:
                os.sched_setaffinity(0, want)
                self.bound = len(want)
        except Exception:
            self.prev = None
        return self

    def __exit__(self, *a):
        if self.prev is not None and self.bound is not None:
            try:
                os.sched_setaffinity(0, self.prev)
            except Exception:
                pass


def workload_name(size: int, streams: int = 2) -> str:
    """One name for the workload, used by both arms (the driver compares the arms on `config`)."""
    return (f"README miniapp (BASELINE configs[1]): dgemm m=n=k={size} NN alpha=1 beta=0, pinned host buffers, tile 5000^3, {streams} streams, "
            "pin_host_buffers=false, copy_c_back=true")


def grid_shape(n: int):
    return {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}.get(n, (1, n))


def run_reference(args, rank: int, world: int) -> None:
    """Reference arm: the UNMODIFIED reference library + cuBLAS (oracle/_ref/libtiledmm_ref.so, built from /root/reference by
    oracle/Makefile) on ONE B200, same buffers / config, through its own public API gpu::gemm."""
    if rank != 0:
        return
    import _util
    size = args.size
    base = {"impl": "reference", "metric": "host-to-host dgemm TFLOP/s", "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(size),
                       "l2": "inputs (2 x 800 MB) larger than L2", "note": "reference is single-GPU: at N>1 rank 0 runs it on one GPU"}}
    try:
        import tiled_mm_b200 as tmm
        ref = _util.Reference(cpu=False)
        with NumaLocal(0):  # same placement policy as our arm
            a = tmm.malloc_pinned(np.float64, size * size); b = tmm.malloc_pinned(np.float64, size * size); c = tmm.malloc_pinned(np.float64, size * size)
            fill_uniform(a, 1); fill_uniform(b, 2)
            np.asarray(c)[:] = 0.0
        ctx = ref.context(np.float64, 2, 5000, 5000, 5000)
        for _ in range(args.warmup):
            ctx.gemm("N", "N", size, size, size, 1.0, a, size, b, size, 0.0, c, size, pin=False, copy_c_back=True)
        with ClockSampler(0) as cs:
            t0 = time.perf_counter()
            for _ in range(args.steps):
                ctx.gemm("N", "N", size, size, size, 1.0, a, size, b, size, 0.0, c, size, pin=False, copy_c_back=True)
            dt = (time.perf_counter() - t0) / args.steps
        ctx.close()
        tf = 2.0 * size**3 / dt * 1e-12
        base.update({"value": round(tf, 3), "ms_per_step": round(dt * 1e3, 3), "clocks": cs.summary(),
                     "cpu_baseline": {"value": round(tf, 3), "unit": "TFLOP/s", "cores": 1, "kind": "reference",
                                      "sample": "unmodified reference Tiled-MM (g++ from /root/reference/src) + cuBLAS 12.9 on one B200; 1 host enqueue thread; "
                                                "the reference has no CPU implementation of this path"},
                     "e2e": {"value": round(tf, 3), "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    except Exception as e:  # the reference .so did not travel / cannot load
        base = {"impl": "reference", "unavailable": f"oracle/_ref/libtiledmm_ref.so not usable: {type(e).__name__}: {e}"[:300]}
    print(json.dumps(base), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--size", type=int, default=10000)
    ap.add_argument("--streams", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3 if args.impl != "reference" else 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import tiled_mm_b200 as tmm

    if not torch.cuda.is_available() or tmm.device_count() < 1:
        raise SystemExit("bench.py needs a B200: tiled_mm_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        os.environ.setdefault("TMM_DIST_TIMEOUT_S", "60")  # a grid call here takes ~60 ms: if a peer dies, give up after a minute instead of the library's 10
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    size = args.size
    m = n = k = size
    flops = 2.0 * m * n * k

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ------------------------------------------------------------------ host buffers (pinned, like gpu::malloc_pinned)
    pr, pc = grid_shape(world)
    gi, gj = rank // pc, rank % pc
    with NumaLocal(local_rank) as numa:  # pinned pages on the NUMA node of this rank's GPU
        a = tmm.malloc_pinned(np.float64, m * k); b = tmm.malloc_pinned(np.float64, k * n); c = tmm.malloc_pinned(np.float64, m * n)
        fill_uniform(a, 100 + gi); fill_uniform(b, 200 + gj)
        np.asarray(c)[:] = 0.0
    ctx = tmm.make_context(np.float64, args.streams, 5000, 5000, 5000)

    # ------------------------------------------------------------------ (1) device-resident kernel throughput -> value, roofline
    st = torch.cuda.current_stream()
    dA = torch.empty(m * k, dtype=torch.float64, device="cuda"); dB = torch.empty(k * n, dtype=torch.float64, device="cuda")
    dC = torch.empty(m * n, dtype=torch.float64, device="cuda")
    dA.copy_(torch.from_numpy(np.asarray(a))); dB.copy_(torch.from_numpy(np.asarray(b)))

    def dev_step():
        tmm.device_gemm(np.float64, "N", "N", m, n, k, 1.0, dA.data_ptr(), m, dB.data_ptr(), k, 0.0, dC.data_ptr(), m, stream=st.cuda_stream)

    for _ in range(args.warmup):
        dev_step()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as cs_dev:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for s0, s1 in evs:
            s0.record(st); dev_step(); s1.record(st)
        e1.record(st)
        barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    launch_ms = sum(s0.elapsed_time(s1) for s0, s1 in evs) / args.steps
    value_tf = world * flops / (dev_ms * 1e-3) * 1e-12
    kernel_tf = flops / (launch_ms * 1e-3) * 1e-12

    # cuBLAS FP64 on the same resident operands (comparator only; never on the product path)
    A2, B2 = dA.view(k, m).t(), dB.view(n, k).t()
    for _ in range(2):
        torch.matmul(A2, B2)
    barrier()
    e0.record(st)
    for _ in range(3):
        torch.matmul(A2, B2)
    e1.record(st); barrier()
    cublas_tf = flops / (e0.elapsed_time(e1) / 3 * 1e-3) * 1e-12
    del A2, B2, dA, dB, dC
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ (2) end to end through the public call -> e2e
    def e2e_step():
        tmm.gemm(ctx, "N", "N", m, n, k, 1.0, a, m, b, k, 0.0, c, m, pin_host_buffers=False, copy_c_back=True)

    if world > 1:
        # rank (gi, gj) owns C block (gi, gj) of the global (pr*m) x (pc*n) x k product: `a` is its A row-panel, `b` its B column-panel.
        # It uploads 1/pc of a and 1/pr of b over its own PCIe link; NCCL all-gathers the shares over NVLink (csrc/tmm_dist.cu).
        from tiled_mm_b200 import multi_gpu
        grid = multi_gpu.GridGemm(ctx, dist)

        def e2e_step():  # noqa: F811
            grid.gemm("N", "N", m, n, k, 1.0, a, m, b, k, 0.0, c, m, pin_host_buffers=False, copy_c_back=True)
    for _ in range(args.warmup):
        e2e_step()
    barrier()
    def timed_e2e():
        launches_before = tmm.total_kernel_launches()
        with ClockSampler(local_rank) as sampler:
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e2e_step()
            barrier()
            ms = (time.perf_counter() - t0) / args.steps * 1e3
        return max_over_ranks(ms), sampler, tmm.total_kernel_launches() - launches_before

    e2e_ms, cs_e2e, launches = timed_e2e()
    # a timed region that saw a hardware / thermal slowdown on any rank is measured again, once (all ranks decide together)
    remeasured = False
    slow = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if max_over_ranks(1.0 if slow & set(cs_e2e.summary()["reasons"]) else 0.0) > 0.0:
        time.sleep(5.0)
        barrier()
        e2e_ms, cs_e2e, launches = timed_e2e()
        remeasured = True
    e2e_tf = world * flops / (e2e_ms * 1e-3) * 1e-12
    stt = ctx.last_stats()
    h2d, d2h, peer = int(stt.h2d_bytes), int(stt.d2h_bytes), int(stt.peer_bytes)
    if world > 1:  # whole-job bytes per step, summed over ranks
        t = torch.tensor([h2d, d2h, peer], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        h2d, d2h, peer = (int(x) for x in t.tolist())
    # sanity: the timed result is a real product (C x = A (B x) on rank 0 at N = 1)
    if True:  # every rank checks its own block (its A row-panel and B column-panel are local)
        x = np.random.default_rng(5).random(n) - 0.5
        lhs = np.asarray(c).reshape(n, m).T @ x
        rhs = np.asarray(a).reshape(k, m).T @ (np.asarray(b).reshape(n, k).T @ x)
        assert np.max(np.abs(lhs - rhs)) <= 1e-15 * k * np.abs(x).sum() + 1e-9, "bench result failed the linearity check"

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_dgemm(size)

    if rank == 0:
        # rooflines for the host-to-host call: min(FP64 peak, AI x PCIe) and the full-duplex variant
        pcie_bytes = 8.0 * (m * k / pc + k * n / pr + m * n)  # per GPU: its upload shares of the shared panels + its C block
        ai = flops / pcie_bytes
        roof_simple = min(FP64_PEAK_TFLOPS, ai * PCIE_H2D_GBS * 1e-3)
        t_duplex = max(flops / (FP64_PEAK_TFLOPS * 1e12), 8.0 * (m * k / pc + k * n / pr) / (PCIE_H2D_GBS * 1e9), 8.0 * m * n / (PCIE_D2H_GBS * 1e9))
        clocks = cs_e2e.summary()
        clocks["remeasured_after_slowdown"] = remeasured
        out = {
            "metric": "host-to-host dgemm TFLOP/s", "value": round(value_tf, 3), "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(dev_ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(size, args.streams) + (f"; weak-scaled over a {pr}x{pc} C-block grid, global {pr*m}x{pc*n}x{k}, A/B panel shares pushed peer-to-peer over NVLink" if world > 1 else ""),
                       "l2": "inputs (A, B = 800 MB each) larger than the 126 MB L2; no flush needed",
                       "host_buffers": f"cudaHostAlloc, first touched on the GPU-local NUMA node ({numa.bound} CPUs)" if numa.bound else "cudaHostAlloc (no NUMA binding applied)",
                       "value_is": "device-resident DGEMM (tmm_device_gemm, operands in HBM)", "e2e_is": "tmm_gemm with host pointers (H2D + GEMM + D2H)"},
            "e2e": {"value": round(e2e_tf, 3), "unit": "TFLOP/s", "ms_per_step": round(e2e_ms, 3),
                    "timing": "host clock around K synchronous calls (each returns only when every stream of the call is idle and host C is complete), "
                              "bracketed by barrier + cudaDeviceSynchronize, max over ranks", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "nvlink_bytes_per_step": peer,
                    "frac_of_host_roofline": round(e2e_tf / world / roof_simple, 4),
                    "host_roofline": {"formula": "min(FP64 peak, AI x PCIe H2D BW)", "tflops": round(roof_simple, 2), "ai_flop_per_byte": round(ai, 1),
                                      "duplex_tflops": round(flops / t_duplex * 1e-12, 2), "active_bound": "fp64" if roof_simple >= FP64_PEAK_TFLOPS - 1e-9 else "pcie",
                                      "pcie_h2d_gbs": PCIE_H2D_GBS, "pcie_d2h_gbs": PCIE_D2H_GBS}},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": round(kernel_tf, 3), "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": round(kernel_tf / FP64_PEAK_TFLOPS, 4),
                         "traffic": 13.2e9, "traffic_note": "dram__bytes_read+write per launch from profiles/r1_ncu_dgemm.md (ncu --set full); algorithmic 2.4e9 B; DRAM at 2.9 % of peak, not the bound",
                         "kernel": "tmm::f64::dgemm_kernel<false,false> (DMMA.8x8x4 fed by TMA)",
                         "peak_source": "FP64 tensor (DMMA) issue peak measured by tools/probe.cu on this pool (profiles/r1_probe_b200.txt); MEASURED_PEAKS.json "
                                        "holds only HBM and bf16 figures, which do not bound an FP64 GEMM",
                         "cublas_dgemm_same_operands_tflops": round(cublas_tf, 3), "algorithmic_flops_per_launch": flops},
            "clocks": clocks, "clocks_device_resident": cs_dev.summary(),
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""CPU suite: the product's REAL scheduler (csrc/tmm_context.cu) and multi-GPU layer (csrc/tmm_dist.cu), compiled as plain C++ and
run over a CPU emulation of the CUDA runtime (tests/emul/): every copy and every GEMM operand is bounds-checked, the GEMM launches
are executed by the oracle, and results must be bit-exact.  This is the only place the 2x4 (8-GPU) grid, both data planes
(peer DMA push with arrival / ack counters, NCCL staging) and the streaming ring with acknowledgements can be exercised without
an 8-GPU box; on hardware the same code was run at 1, 2 and 4 GPUs (profiles/r1_final_*).  Underneath runs a vector-clock
happens-before race detector over all device-memory accesses (it found a write-after-read hazard on the ring slot's own share in
the streaming grid path, fixed in csrc/tmm_dist.cu `slot_pushed`).  Test infrastructure only."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EMUL = ROOT / "tests" / "emul"


@pytest.fixture(scope="module")
def emul_build():
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(["make", "-C", str(EMUL), "-j4", "all"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    return EMUL / "_build"


def _worker(build, args, devices, extra_env=None, expect_failure=None):
    env = dict(os.environ)
    env.update({"TMM_EMUL_DEVICES": str(devices), "TMM_EMUL_MEM_MB": "2048", "TMM_NCCL_LIB": str(build / "libnccl.so.2"), "TMM_DIST_TIMEOUT_S": "60"})
    env.pop("TMM_DIST_NCCL", None)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "_emul_worker.py"), *map(str, args)], capture_output=True, text=True, env=env, timeout=900)
    if expect_failure:
        assert r.returncode != 0 and expect_failure in r.stderr, (r.stdout[-2000:], r.stderr[-4000:])
        return
    assert r.returncode == 0 and "EMUL_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


def test_single_gpu_scheduler_on_emulated_runtime(emul_build):
    """resident and streaming regimes, all op pairs, four types, degenerate shapes, pageable buffers; no out-of-bounds access, no
    panel outside the TMA contract, no async copy from pageable memory, no device memory left behind"""
    _worker(emul_build, ["single"], 1)


def test_race_detector_catches_a_schedule_without_event_waits(emul_build):
    """Mutation check of the happens-before detector underneath all of these tests: with every cudaStreamWaitEvent ignored, the very
    same schedule must be reported as racy (H2D -> GEMM -> D2H dependencies are gone)."""
    _worker(emul_build, ["single"], 1, {"TMM_EMUL_DROP_WAITS": "1"}, expect_failure="unordered conflicting accesses")


@pytest.mark.parametrize("devices", [2, 3, 4, 6, 8])
def test_gpu_grid_dma_push_on_emulated_runtime(emul_build, devices):
    """one process, `devices` emulated GPUs, shares pushed into mapped peer panels with arrival / ack counters"""
    _worker(emul_build, ["grid", devices, "direct"], devices)


@pytest.mark.parametrize("devices", [2, 4, 8])
def test_gpu_grid_nccl_staging_on_emulated_runtime(emul_build, devices):
    """same, with the NCCL all-gather staging ring as data plane (TMM_DIST_NCCL=1)"""
    _worker(emul_build, ["grid", devices, "nccl"], devices, {"TMM_DIST_NCCL": "1"})


@pytest.mark.parametrize("devices", [1, 2, 4, 8])
def test_full_size_configs_dry_run(emul_build, devices):
    """BASELINE configs[3] (zgemm 20000 x 20000 x 500000) and configs[4] (dgemm 100000^3, 240 GB) plus a C that exceeds one HBM, walked
    through the real scheduler on `devices` emulated 180 GB GPUs with address-only memory: every copy / launch is bounds- and
    order-checked with its real 64-bit offsets, the exchange protocol must make progress, and each A / B / C element must cross the
    host link exactly once whenever the per-GPU share fits (resident regime)."""
    _worker(emul_build, ["dry", devices], devices, {"TMM_EMUL_DRY": "1", "TMM_EMUL_MEM_MB": "182000"})


@pytest.mark.parametrize("devices,cases,seed,plane", [(1, 600, 11, "direct"), (2, 80, 12, "direct"), (4, 80, 13, "direct"), (8, 80, 14, "direct"), (6, 60, 15, "nccl")])
def test_randomised_sweep_on_emulated_runtime(emul_build, devices, cases, seed, plane):
    """random types / ops / shapes / lds / scalars / tile hints / budgets / copy modes on reused contexts, bit-exact against the oracle,
    with bounds, TMA-contract and race checks on (the way SURVEY pinned the reference's own valid domain, turned on this library)"""
    _worker(emul_build, ["sweep", devices, cases, seed], devices, {"TMM_DIST_NCCL": "1"} if plane == "nccl" else None)


def test_replan_when_free_memory_shrinks_between_calls(emul_build):
    """a device allocation that fails before anything is enqueued is recovered by releasing staging storage and re-planning"""
    _worker(emul_build, ["replan"], 1, {"TMM_EMUL_MEM_MB": "32"})


@pytest.mark.parametrize("stripes,devices", [(2, 1), (3, 1), (4, 1), (4, 4)])
def test_phase1_column_stripes_with_staggered_c_upload(emul_build, stripes, devices):
    """Phase 1 cut into several column stripes (forced with TMM_PLAN_P1SPLIT; the shapes of this suite are too small to get them
    by themselves), in round 1's C-first order (TMM_PLAN_DEFER_C=0, kept as a switch): with beta != 0 stripe s's share of C is uploaded
    right before k-chunk s and the stripe catches up on the chunks that arrived earlier - including the case of fewer chunks than stripes."""
    _worker(emul_build, ["sweep", devices, 400 if devices == 1 else 100, 40 + stripes], devices, {"TMM_PLAN_P1SPLIT": str(stripes), "TMM_PLAN_DEFER_C": "0"})


@pytest.mark.parametrize("stripes,devices", [(1, 1), (3, 1), (4, 4), (4, 8)])
def test_beta_times_c_added_after_the_accumulation(emul_build, stripes, devices):
    """The default order for beta != 0 (round 2): every block accumulates from zero, the caller's C travels behind the A / B panels into a staging
    copy and C += beta * C_host runs when the block's last launch and its share of the copy are both done; the first phase-2 block's B is
    fetched ahead of the C stripes.  Bit-exact against the oracle on integer data, race detector and bounds checks on, one GPU and grids."""
    _worker(emul_build, ["sweep", devices, 400 if devices == 1 else 100, 70 + stripes], devices, {"TMM_PLAN_P1SPLIT": str(stripes)})


@pytest.mark.parametrize("stripes,devices,c_first", [(1, 1, False), (3, 1, False), (4, 1, True), (2, 4, False), (4, 8, False)])
def test_complex_float_operands_prepared_once(emul_build, stripes, devices, c_first):
    """complex<float> on the tensor-core embedding (round 2): the scheduler embeds every k-chunk of A once into the context's A' buffer (one
    stream, the stripe streams wait on its event), phase 2 multiplies the A' the chunks add up to, op(B) = T / C pieces are split by the launch
    that consumes them; a sub-block off the 16-byte grid falls back to a self-contained launch.  Stand-ins with the device passes' layouts and
    declared accesses (tests/emul/emul_blas.cpp): bit-exact against the oracle's complex GEMM on integer data, range checker and race detector on."""
    env = {"TMM_PLAN_P1SPLIT": str(stripes), "TMM_EMUL_C32_TC": "1", "TMM_EMUL_DTYPES": "c"}
    if c_first:
        env["TMM_PLAN_DEFER_C"] = "0"
    _worker(emul_build, ["sweep", devices, 300 if devices == 1 else 80, 90 + stripes], devices, env)


def test_race_detector_sees_the_prepared_operands(emul_build):
    """Mutation check for the test above: with the event waits dropped, a stripe stream multiplies a chunk of A' that another stream is still writing."""
    _worker(emul_build, ["sweep", 1, 60, 93], 1, {"TMM_PLAN_P1SPLIT": "3", "TMM_EMUL_C32_TC": "1", "TMM_EMUL_DTYPES": "c", "TMM_EMUL_DROP_WAITS": "1"},
            expect_failure="unordered conflicting accesses")


def test_one_c_stripe_per_k_chunk_experiment(emul_build):
    """TMM_PLAN_CSTRIPES=chunks (schedule experiment of the C-first order: as many column stripes as k-chunks, more stripes than streams)"""
    _worker(emul_build, ["sweep", 1, 500, 61], 1, {"TMM_PLAN_CSTRIPES": "chunks", "TMM_PLAN_DEFER_C": "0"})


def test_scheduler_under_address_and_ub_sanitizers():
    """The same scheduler sources built with -fsanitize=address,undefined (the reference has no sanitizer coverage): a random
    sweep on one device and on a 2x2 grid must finish without a report (heap / stack overflows in the host-side bookkeeping, signed
    overflow in the 64-bit offset arithmetic, ...)."""
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not asan or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    r = subprocess.run(["make", "-C", str(EMUL), "-j4", "SAN=1", "all"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    build = EMUL / "_build_asan"
    san_env = {"TMM_EMUL_LIB": str(build / "libtiledmm_emul.so"), "LD_PRELOAD": asan, "ASAN_OPTIONS": "detect_leaks=0:halt_on_error=1",
               "UBSAN_OPTIONS": "halt_on_error=1:print_stacktrace=1"}
    _worker(build, ["sweep", 1, 250, 81], 1, san_env)
    _worker(build, ["sweep", 4, 60, 82], 4, dict(san_env, TMM_PLAN_P1SPLIT="3"))
    _worker(build, ["dry", 8], 8, dict(san_env, TMM_EMUL_DRY="1", TMM_EMUL_MEM_MB="182000"))


def test_one_context_per_host_thread_concurrently(emul_build):
    """eight host threads, one context each, calls in flight at the same time on one device"""
    _worker(emul_build, ["threads", 8, 60], 1)


@pytest.mark.parametrize("devices,cases", [(1, 2000), (4, 200), (8, 200)])
def test_address_only_sweep_over_realistic_sizes(emul_build, devices, cases):
    """random shapes up to 60000 per dimension, all types / ops, budgets from 1 GiB to all of HBM, walked without arithmetic: every plan
    structure the scheduler produces at scale is bounds-, order- and byte-count-checked"""
    _worker(emul_build, ["drysweep", devices, cases, 500 + devices], devices, {"TMM_EMUL_DRY": "1", "TMM_EMUL_MEM_MB": "182000"})


def test_tmm_devices_environment_switch(emul_build):
    """TMM_DEVICES=4: an unchanged caller gets a 2x2 grid; device-resident C calls still work (first device)"""
    _worker(emul_build, ["auto", 4], 4, {"TMM_DEVICES": "4"})


def test_fault_injection_allocation_and_copy_failures(emul_build):
    """the n-th cudaMalloc / cudaMemcpy2DAsync of a call fails: error or recovery, never a crash, a hang or a leak; the context stays usable"""
    _worker(emul_build, ["faults"], 1)


@pytest.mark.parametrize("devices,plane", [(2, "direct"), (8, "direct"), (4, "nccl")])
def test_grid_gives_a_call_up_together_when_one_gpu_cannot_allocate(emul_build, devices, plane):
    """fault injection on a GPU grid: an allocation fails on one rank - all ranks return an error for that call (no rank is left waiting
    for shares), nothing leaks, and the same contexts compute the next call correctly"""
    _worker(emul_build, ["gridfaults", devices], devices, {"TMM_DIST_NCCL": "1"} if plane == "nccl" else None)


@pytest.mark.parametrize("world,extra", [(2, {"TMM_DIST_FORCE_IPC": "1"}), (4, {"TMM_DIST_FORCE_IPC": "1"}), (8, {"TMM_DIST_FORCE_IPC": "1"}), (4, {"TMM_DIST_NCCL": "1"}), (8, {})])
def test_one_rank_per_gpu_entry_point_with_emulated_ipc(emul_build, world, extra):
    """tmm_context_attach_grid (the torchrun / MPI entry point) played by one host thread per emulated GPU: growing, shrinking and
    streaming calls on reused contexts, bit-exact per block; through the emulated CUDA IPC calls the import bookkeeping is checked
    (re-import when a peer's buffer changed, outgrown buffers retired instead of freed while imported, everything closed at teardown)."""
    _worker(emul_build, ["ranks", world], world, extra)


@pytest.mark.parametrize("world,extra", [(2, {}), (8, {}), (4, {"TMM_DIST_BOARD": "0"})])
def test_rank_that_fails_before_the_agreement_round_takes_the_grid_with_it(emul_build, world, extra):
    """A rank that rejects its arguments before anything collective (here: ld_a too small on one rank only) still plays the call's agreement
    round and reports the failure: every rank returns an error for that call instead of waiting for the missing rank, and the next call on
    the same contexts is bit-exact.  Run on the shared-memory control board (default) and on the NCCL control collectives (TMM_DIST_BOARD=0)."""
    _worker(emul_build, ["rankfails", world], world, extra)


@pytest.mark.parametrize("world", [4])
def test_grid_entry_point_on_nccl_control_collectives(emul_build, world):
    """the control plane without the shared-memory board (what a box without /dev/shm falls back to)"""
    _worker(emul_build, ["ranks", world], world, {"TMM_DIST_BOARD": "0", "TMM_DIST_FORCE_IPC": "1"})


@pytest.mark.parametrize("devices,mode", [(2, "grid"), (4, "sweep"), (8, "sweep"), (8, "ranks")])
def test_upload_shares_follow_the_measured_link_rates(emul_build, devices, mode):
    """The ranks of a link do not sit behind equally fast host links; with measured rates (made up per device on the emulated runtime,
    TMM_EMUL_LINK_RATES=1) the shares of a shared panel are unequal - boundaries agreed by integer arithmetic on every rank, every element
    still uploaded exactly once (the byte accounting of the workers checks it), results bit-exact, no race, on resident and streaming calls."""
    env = {"TMM_EMUL_LINK_RATES": "1", "TMM_DEBUG": "0"}
    if mode == "grid":
        _worker(emul_build, ["grid", devices, "direct"], devices, env)
    elif mode == "ranks":
        _worker(emul_build, ["ranks", devices], devices, dict(env, TMM_DIST_FORCE_IPC="1"))
    else:
        _worker(emul_build, ["sweep", devices, 120, 90 + devices], devices, env)

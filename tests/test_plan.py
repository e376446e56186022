"""CPU suite: the scheduler's host logic (tmm_plan.cpp) through tmm_plan_describe - no GPU involved."""
import numpy as np
import pytest

GB = 1 << 30
SHAPES = [(1000, 1000, 1000), (10000, 10000, 10000), (1234, 4567, 1357), (12345, 23456, 67891), (50, 200, 21), (5, 2, 2), (3001, 2003, 4099),
          (20000, 20000, 500000), (100000, 100000, 100000), (100000, 1000, 10000), (1000, 100000, 777), (1, 1, 1), (129, 65, 100000)]


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.complex64, np.complex128])
@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("tt", ["NN", "TN", "NT", "CC"])
def test_plan_invariants(tmm, dtype, shape, tt):
    m, n, k = shape
    es = np.dtype(dtype).itemsize
    for budget in (150 * GB, 2 * GB, 256 << 20):
        for copy_c_back in (True, False):
            p = tmm.plan_describe(dtype, tt[0], tt[1], m, n, k, True, copy_c_back, budget)
            if p["error"]:
                # only legal when even the smallest ring cannot fit
                assert budget <= 2 * GB
                continue
            align = 128 // es
            for key in ("pitch_a", "pitch_b"):
                assert p[key] % align == 0
            assert p["pitch_c"] == (m if not copy_c_back else -(-m // align) * align)
            assert p["bytes_a"] + p["bytes_b"] + p["bytes_c"] <= budget
            if p["regime"] == 0:
                assert sum(p["chunks"]) == k and all(c > 0 for c in p["chunks"])
                assert p["n1"] + sum(p["blocks"]) == n and all(b > 0 for b in p["blocks"])
                assert 0 < p["n1"] <= n
                # sub-panel origins stay 16-byte aligned for TMA: chunk and block starts are multiples of 64 elements
                starts = np.cumsum([0] + p["chunks"][:-1])
                assert all(s % 64 == 0 for s in starts)
                cols = np.cumsum([p["n1"]] + p["blocks"][:-1]) if p["blocks"] else []
                assert all(c % 64 == 0 for c in cols)
                # every element crosses PCIe exactly once
                assert p["h2d_bytes"] == es * (m * k + k * n + m * n)
            else:
                assert p["kc"] % 64 == 0 or p["kc"] >= k
                assert p["MB"] >= 1 and p["NB"] >= 1 and p["slots"] >= 2
                if p["MB"] < m:
                    assert p["MB"] % 128 == 0
                if p["NB"] < n:
                    assert p["NB"] % 64 == 0
                bm, bn = -(-m // p["MB"]), -(-n // p["NB"])
                assert p["h2d_bytes"] == es * (m * k * bn + k * n * bm + m * n)
                assert p["c_is_full"] == (not copy_c_back)
            assert p["d2h_bytes"] == (es * m * n if copy_c_back else 0)


def test_headline_config_plan(tmm):
    """BASELINE config 2 (dgemm 10000^3): resident, A and B cross PCIe once (the reference moves 3.2 GB)."""
    p = tmm.plan_describe(np.float64, "N", "N", 10000, 10000, 10000, False, True, 150 * GB)
    assert p["regime"] == 0
    assert p["h2d_bytes"] == 1_600_000_000 and p["d2h_bytes"] == 800_000_000
    assert p["chunks"][0] <= 512            # short prologue
    assert p["blocks"][-1] <= 512           # short D2H tail
    assert p["launches"] <= 16


def test_out_of_core_configs_stream(tmm):
    p4 = tmm.plan_describe(np.complex128, "N", "N", 20000, 20000, 500000, False, True, 150 * GB)   # BASELINE config 4
    assert p4["regime"] == 1 and p4["MB"] == 20000 and p4["NB"] == 20000 and p4["n_cbuf"] == 1
    p5 = tmm.plan_describe(np.float64, "N", "N", 100000, 100000, 100000, False, True, 150 * GB)    # BASELINE config 5 on one GPU
    assert p5["regime"] == 1
    assert p5["h2d_bytes"] >= 8 * 2 * 100000**2
    # tiny budget forces C super-blocks with two buffers
    p = tmm.plan_describe(np.float64, "N", "N", 6000, 6000, 3000, False, True, 128 << 20)
    assert p["regime"] == 1 and p["n_cbuf"] == 2 and (p["MB"] < 6000 or p["NB"] < 6000)


def test_headline_plan_is_the_measured_one(tmm):
    """The plan of BASELINE configs[1] is pinned to the one that was measured on hardware (profiles/r2_sweep_plan.txt, r2_sweep_plan_fine2.txt:
    chunk growth 1.52 from a 192-wide first chunk, 9 chunks, 58.02 - 58.06 ms against 58.57 ms for round 1's 12-chunk plan): planner changes
    made without a GPU must not move it."""
    p = tmm.plan_describe(np.float64, "N", "N", 10000, 10000, 10000, False, True, int(0.92 * 178e9))
    assert p["regime"] == 0 and p["n1"] == 5504
    assert p["chunks"] == [192, 256, 384, 576, 832, 1216, 1792, 2048, 2704]
    assert p["blocks"] == [1664, 1664, 896, 272]


def test_timeline_model_reproduces_the_measured_calls(tmm):
    """tools/model_resident.py is only a design aid, but the schedule changes made after the last hardware run lean on it: it has to
    reproduce the two calls it was calibrated against (dgemm 10000^3: 58.60 ms, sgemm 10000^3 with the round-1 plan: 21.73 ms)."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
    import model_resident as mr
    t, _, _, _ = mr.predict(10000, 10000, 10000, 0.0)
    assert abs(t * 1e3 - 58.6) < 1.0, t
    round1_f32_plan = {"regime": 0, "n1": 2368, "chunks": [256, 320, 384, 512, 640, 832, 1088, 1408, 1856, 2704], "blocks": [1664, 1664, 1664, 1664, 704, 272]}
    saved = mr.P
    try:
        mr.P = 140e12
        t, _, _, _ = mr.predict(10000, 10000, 10000, 0.0, plan=round1_f32_plan, es=4)
    finally:
        mr.P = saved
    assert abs(t * 1e3 - 21.73) < 1.0, t

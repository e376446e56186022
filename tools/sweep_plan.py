#!/usr/bin/env python
"""Schedule knobs of the resident regime swept on ONE set of host buffers (GPU box; development tool): which first chunk, chunk growth,
phase-1 stripe count and column-block width give the shortest host-to-host dgemm 10000^3?   python tools/sweep_plan.py [--reps 6]"""
import argparse, itertools, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10000); ap.add_argument("--reps", type=int, default=6); ap.add_argument("--beta", type=float, default=0.0)
ap.add_argument("--quick", action="store_true")
args = ap.parse_args()
import tiled_mm_b200 as tmm
n = args.n
a = tmm.malloc_pinned(np.float64, n * n); b = tmm.malloc_pinned(np.float64, n * n); c = tmm.malloc_pinned(np.float64, n * n)
rng = np.random.default_rng(0)
for arr in (a, b):
    for off in range(0, arr.size, 1 << 24):
        arr[off:off + (1 << 24)] = rng.random(min(1 << 24, arr.size - off)) - 0.5
ctx = tmm.make_context(np.float64, 2, 5000, 5000, 5000)
KNOBS = ["TMM_PLAN_KC0", "TMM_PLAN_GROWTH", "TMM_PLAN_KCMAX", "TMM_PLAN_P1SPLIT", "TMM_PLAN_NB", "TMM_PLAN_MARGIN"]

def run(cfg):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in cfg.items():
        os.environ[k] = str(v)
    ts = []
    for _ in range(args.reps):
        t0 = time.perf_counter()
        tmm.gemm(ctx, "N", "N", n, n, n, 1.0, a, n, b, n, args.beta, c, n, pin_host_buffers=False, copy_c_back=True)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    st = ctx.last_stats()
    return ts[0], ts[len(ts) // 2], st.k_chunks, st.c_blocks

run({})
base = run({})
print(f"default: best {base[0]:.2f} median {base[1]:.2f} ms  chunks {base[2]} blocks {base[3]}", flush=True)
if os.environ.get("SWEEP_FINE") == "1":   # second pass of round 2: around the adopted growth (1.15 r = 1.52 at 10000^3)
    fine = [{"TMM_PLAN_KC0": k0} for k0 in (64, 96, 128, 160, 192, 224)] + [{"TMM_PLAN_GROWTH": 1.45, "TMM_PLAN_KC0": k0} for k0 in (128, 192)] + \
           [{"TMM_PLAN_GROWTH": g} for g in (1.3, 1.38, 1.45, 1.6, 1.7)] + [{"TMM_PLAN_GROWTH": 1.52, "TMM_PLAN_KC0": k0} for k0 in (192, 320, 384, 512)] + \
           [{"TMM_PLAN_GROWTH": 1.4, "TMM_PLAN_KC0": 384}, {"TMM_PLAN_GROWTH": 1.4, "TMM_PLAN_KC0": 512}, {"TMM_PLAN_KCMAX": 3072, "TMM_PLAN_GROWTH": 1.52}, {"TMM_PLAN_P1SPLIT": 3}, {"TMM_PLAN_P1SPLIT": 2}]
    for cfg in fine:
        r = run(cfg)
        print(f"{cfg}: best {r[0]:.2f} median {r[1]:.2f} ms  chunks {r[2]} blocks {r[3]}", flush=True)
    r = run({})
    print(f"default again: best {r[0]:.2f} median {r[1]:.2f} ms", flush=True)
    ctx.close()
    sys.exit(0)
one_at_a_time = [{"TMM_PLAN_KC0": v} for v in (128, 192, 384)] + [{"TMM_PLAN_GROWTH": v} for v in (1.25, 1.5, 2.0)] + \
                [{"TMM_PLAN_KCMAX": v} for v in (1024, 3072, 4096)] + [{"TMM_PLAN_P1SPLIT": v} for v in (1, 2, 3)] + \
                [{"TMM_PLAN_NB": v} for v in (1024, 1536, 3072)] + [{"TMM_PLAN_MARGIN": v} for v in (1.1, 1.2, 1.5)]
results = []
for cfg in one_at_a_time:
    r = run(cfg)
    results.append((r[0], r[1], cfg, r[2], r[3]))
    print(f"{cfg}: best {r[0]:.2f} median {r[1]:.2f} ms  chunks {r[2]} blocks {r[3]}", flush=True)
if not args.quick:
    # combine the best value of every knob that helped
    best_per_knob = {}
    for best, med, cfg, *_ in results:
        (k, v), = cfg.items()
        if med < base[1] - 0.05 and (k not in best_per_knob or med < best_per_knob[k][0]):
            best_per_knob[k] = (med, v)
    combo = {k: v for k, (_, v) in best_per_knob.items()}
    if len(combo) > 1:
        r = run(combo)
        print(f"combined {combo}: best {r[0]:.2f} median {r[1]:.2f} ms  chunks {r[2]} blocks {r[3]}", flush=True)
    r = run({})
    print(f"default again: best {r[0]:.2f} median {r[1]:.2f} ms", flush=True)
ctx.close()

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export TMM_DIST_TIMEOUT_S=30
echo "== single process e2e 2 devices =="; timeout 100 python tools/e2e.py --devices 2 --m 10000 --n 20000 --k 10000 --reps 4 2>&1 | tail -6 | tee gpurun_out/dbg_single.txt
echo "== pytest multi =="; timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q --timeout 120 2>&1 | tail -5

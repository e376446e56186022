#!/bin/bash
# Quick GPU pass: parity suite + smoke + kernel bench + short e2e bench.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
echo "== pytest gpu =="; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== smoke =="; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee gpurun_out/smoke.txt
echo "== kbench =="; timeout 600 python tools/kbench.py ${KB:-all} 2>&1 | tee gpurun_out/kbench.txt
echo "== bench ours =="; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.json

#!/bin/bash
# Round 2, final single-GPU run at HEAD: a last schedule knob, full GPU suite, smoke, both bench arms.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
nvidia-smi -L
echo "##### width of the final column block (its D2H is the exposed tail): default, 128, 64"
for t in 0 128 64; do TMM_PLAN_TAIL=$t timeout 90 python tools/e2e.py --reps 8 2>&1 | tail -1; done
echo "##### pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
echo "##### smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
echo "##### bench.py --impl reference"; timeout 300 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1
echo "##### bench.py (ours)"; timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r2_final2.txt

#!/bin/bash
# Round 2, final single-GPU run at HEAD: full GPU suite, smoke, both bench arms.
# (The run recorded in profiles/r2_final_single_gpu_2.txt also tried a knob that carved a narrow last column block off the plan - no effect, removed.)
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
nvidia-smi -L
timeout 90 python tools/e2e.py --reps 8 2>&1 | tail -1
echo "##### pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
echo "##### smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
echo "##### bench.py --impl reference"; timeout 300 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1
echo "##### bench.py (ours)"; timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r2_final2.txt

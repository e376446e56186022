#!/bin/bash
# Round 2, one B200: full GPU suite, the bench line of both arms, optional extras.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
nvidia-smi -L; nproc; free -g | head -2 | tail -1
[ "$1" = "pin" ] && { echo "##### pin_probe"; timeout 300 ./build/pin_probe 8; }
echo "##### pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -8
echo "##### smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
echo "##### bench.py (ours)"; timeout 300 python bench.py --steps 10 --warmup 3 2>&1 | tail -1
echo "##### bench.py --impl reference"; timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r2_single.txt

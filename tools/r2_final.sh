#!/bin/bash
# Round 2, final single-GPU run at HEAD: SGEMM window experiment, full GPU suite, smoke, both bench arms, published sweep refresh.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
nvidia-smi -L
echo "##### tcgen05 SGEMM: k-blocks per TMEM accumulation window (default 4) - does the plain-TF32 mode want longer windows?"
for w in 4 16 64; do echo "window $w"; TMM_TC_WINDOW=$w timeout 60 ./build/tc_test benchone N N 8192 8192 8192 0; done
echo "##### pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
echo "##### smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
echo "##### bench.py --impl reference"; timeout 300 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1
echo "##### bench.py (ours)"; timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1
echo "##### published experiment (alpha = beta = 1), both arms"; timeout 500 python tools/sweep_published.py --reps 2 --sizes 4000,8000,10000,12000,16000,20000,24000,28000,32000 2>&1 | tail -11
echo "##### device-resident C (copy_c_back = false), beta = 0"; timeout 300 python tools/sweep_published.py --reps 2 --beta 0 --copy-c-back 0 --sizes 4000,10000,16000 2>&1 | tail -4
} 2>&1 | tee gpurun_out/r2_final.txt

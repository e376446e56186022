#!/bin/bash
# Round 2, call 1 (one B200): everything written after round 1's GPU budget ran out gets its first run; results decide promote / delete.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
nvidia-smi -L; nproc; free -g | head -2
echo "##### experimental pytest (cgemm tc, bf16, device pointers)"
TMM_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -m gpu -q --timeout 120 -k "not tmem and not int8" 2>&1 | tail -25
echo "##### tc_test cgemm"; timeout 200 ./build/tc_test cgemm 2>&1 | grep -v " OK$" | tail -30
echo "##### native bf16"; TMM_BF16_NATIVE=1 TMM_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_experimental_gpu.py -m gpu -q -k bf16 --timeout 100 2>&1 | tail -6
} 2>&1 | tee gpurun_out/r2_call1_a.txt
timeout 400 bash tools/gpu_round2_tc.sh > /dev/null 2>&1
timeout 400 bash tools/gpu_round2_i8.sh > /dev/null 2>&1
{
echo "##### tf32 trunc split"
TMM_TC_SPLIT=trunc timeout 120 ./build/tc_test precision 2>&1 | grep -E "precision|tmm fp32" | head -20
TMM_TC_SPLIT=trunc timeout 60 ./build/tc_test benchone N N 8192 8192 8192 0
echo "##### e2e baseline 10000^3 + trace"
timeout 60 python tools/e2e.py --reps 6 2>&1 | tail -2
TMM_TRACE=1 timeout 60 python tools/e2e.py --reps 2 2>&1 | tail -90
} 2>&1 | tee gpurun_out/r2_call1_b.txt

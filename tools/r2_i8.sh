#!/bin/bash
# Round 2: the opt-in FP64 emulation after the slicing rework and the slice cache in the resident schedule (one B200).
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
nvidia-smi -L
echo "##### parity: int8 tests (8 and 7 slices) + a scheduler sweep with the cache on"
timeout 300 python -m pytest tests/test_experimental_gpu.py -m gpu -q -k int8 --timeout 200 2>&1 | tail -4
TMM_F64_MATH=i8:7 timeout 500 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 300 2>&1 | tail -12
echo "##### device-resident 10000^3: DMMA, 7 slices, 8 slices"
timeout 60 ./build/devtest benchone N N 10000 10000 10000 0
TMM_F64_MATH=i8:7 timeout 90 ./build/devtest benchone N N 10000 10000 10000 0
TMM_F64_MATH=i8:7 timeout 90 ./build/devtest benchone T T 10000 10000 10000 0
TMM_F64_MATH=i8 timeout 90 ./build/devtest benchone N N 10000 10000 10000 0
TMM_F64_MATH=i8:7 timeout 90 ./build/devtest benchone N N 10000 1408 512 1
echo "##### host to host 10000^3: DMMA, 7 slices (cache), 8 slices"
timeout 90 python tools/e2e.py --reps 5 2>&1 | tail -1
TMM_F64_MATH=i8:7 timeout 90 python tools/e2e.py --reps 5 2>&1 | tail -3
TMM_F64_MATH=i8 timeout 90 python tools/e2e.py --reps 5 2>&1 | tail -1
TMM_F64_MATH=i8:7 timeout 90 python tools/e2e.py --reps 4 --beta 1 2>&1 | tail -1
TMM_F64_MATH=i8:7 timeout 90 python tools/e2e.py --reps 4 --tt TN 2>&1 | tail -1
TMM_F64_MATH=i8:7 timeout 90 python tools/e2e.py --reps 4 --m 16000 --n 16000 --k 16000 2>&1 | tail -1
echo "##### trace, 7 slices"
TMM_F64_MATH=i8:7 TMM_TRACE=1 timeout 90 python tools/e2e.py --reps 2 2>&1 | tail -75
} 2>&1 | tee gpurun_out/r2_i8.txt

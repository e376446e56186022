#!/bin/bash
# Round 2: the opt-in FP64 emulation (TMM_F64_MATH=i8[:S]): parity, device-resident and host-to-host timings.  (one B200)
# (The run recorded in profiles/r2_f64_i8_e2e.txt also compared a 2 x 2-cluster TMA-multicast variant of the kernel, since removed:
#  profiles/r2_i8_cluster_experiment.txt.)
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
nvidia-smi -L
echo "##### parity: int8 tests (8 and 7 slices, device GEMM and through the scheduler)"
timeout 400 python -m pytest tests/test_experimental_gpu.py -m gpu -q -k int8 --timeout 200 2>&1 | tail -6
echo "##### device-resident 10000^3, 7 slices; other ops and shapes"
TMM_F64_MATH=i8:7 timeout 90 ./build/devtest benchone N N 10000 10000 10000 0
TMM_F64_MATH=i8:7 timeout 90 ./build/devtest benchone T T 10000 10000 10000 0
TMM_F64_MATH=i8 timeout 90 ./build/devtest benchone N N 10000 10000 10000 0
TMM_F64_MATH=i8:7 timeout 90 ./build/devtest benchone N N 10000 1408 512 1
TMM_F64_MATH=i8:7 timeout 90 ./build/devtest benchone N N 16384 16384 16384 0
echo "##### check vs cuBLAS"; TMM_F64_MATH=i8:7 timeout 120 ./build/devtest check 2>&1 | grep -c " OK"; TMM_F64_MATH=i8:7 timeout 120 ./build/devtest check 2>&1 | grep -E "FAIL|rror" | head -5
echo "##### host to host 10000^3, 7 slices; DMMA for reference"
TMM_F64_MATH=i8:7 timeout 90 python tools/e2e.py --reps 6 2>&1 | tail -4
timeout 90 python tools/e2e.py --reps 5 2>&1 | tail -1
TMM_F64_MATH=i8:7 timeout 90 python tools/e2e.py --reps 4 --beta 1 2>&1 | tail -1
TMM_F64_MATH=i8:7 timeout 90 python tools/e2e.py --reps 4 --m 16000 --n 16000 --k 16000 2>&1 | tail -1
TMM_F64_MATH=i8:7 TMM_PLAN_TAPER=0 timeout 90 python tools/e2e.py --reps 6 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r2_i8.txt

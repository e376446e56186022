#!/bin/bash
# Round 2, experimental FP64 emulation on the int8 tensor cores (TMM_F64_MATH=i8[:slices], gemm_f64_i8.cu; one B200, ~4 min):
#   make -C tools && gpurun --timeout 420 -- 'bash tools/gpu_round2_i8.sh'
# Default stays DMMA.  The kernel's waits are guarded (a broken pipeline traps after ~10 s).  What to look at:
#   1. devtest check: max|diff| / k against cuBLAS DGEMM for all four op pairs (bound of the parity tests: 1e-15)
#   2. device-resident rate at 10000^3 / 8192^3 for 8 and 7 slices against the DMMA kernel (35.4 TF) -> is the kind::i8 pipeline efficient enough?
#   3. host-to-host 10000^3: does the call become PCIe-bound (46 TF roofline instead of 36.9)?
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
D=./build/devtest
{
nvidia-smi -L | head -1
echo "== unit probe: one kind::i8 MMA on hand-written tiles =="; timeout 60 ./build/tc_probe2 4; timeout 60 ./build/tc_probe2 7   # 7 = the CTA-pair shape of the i8p version
echo "== check vs cuBLAS, 8 slices =="; TMM_F64_MATH=i8 timeout 120 $D check 2>&1 | tail -45
echo "== check vs cuBLAS, 7 slices =="; TMM_F64_MATH=i8:7 timeout 120 $D check 2>&1 | grep -E "FAIL|error|rc=" | head -20
echo "== device-resident, DMMA (reference point) =="; timeout 60 $D benchone N N 10000 10000 10000 0
for s in 8 7 6; do echo "== device-resident, int8 emulation, $s slices =="; TMM_F64_MATH=i8:$s timeout 90 $D benchone N N 10000 10000 10000 0; done
echo "== CTA-pair version (i8p: 256 x 256 tiles, cta_group::2, one launch per group) =="
TMM_F64_MATH=i8p timeout 120 $D check 2>&1 | grep -E "FAIL|error|rc=|OK" | tail -24
for s in 8 7; do TMM_F64_MATH=i8p:$s timeout 90 $D benchone N N 10000 10000 10000 0; done
TMM_F64_MATH=i8p timeout 90 python tools/e2e.py --reps 4 2>&1 | tail -1
echo "== other shapes =="
TMM_F64_MATH=i8 timeout 90 $D benchone T T 8192 8192 8192 0
TMM_F64_MATH=i8 timeout 90 $D benchone N N 10000 1408 512 1     # a phase-1 launch shape of the scheduler: slicing overhead shows here
echo "== host-to-host 10000^3: DMMA, then 8 and 7 slices =="
timeout 60 python tools/e2e.py --reps 4 2>&1 | tail -1
TMM_F64_MATH=i8 timeout 90 python tools/e2e.py --reps 4 2>&1 | tail -1
TMM_F64_MATH=i8:7 timeout 90 python tools/e2e.py --reps 4 2>&1 | tail -1
TMM_F64_MATH=i8p TMM_PLAN_F64_FLOPS=60e12 timeout 90 python tools/e2e.py --reps 4 2>&1 | tail -1   # schedule sized for a faster GEMM (wider first block)
TMM_F64_MATH=i8 TMM_PLAN_P1SPLIT=1 timeout 90 python tools/e2e.py --reps 4 2>&1 | tail -1   # one phase-1 stripe: every A chunk is sliced once instead of four times
echo "== gated pytest (one process per variant) =="
for v in "8-slices and not pairs" "7-slices and not pairs" "pairs-8-slices" "pairs-7-slices"; do TMM_EXPERIMENTAL=1 timeout 150 python -m pytest tests/test_experimental_gpu.py -m gpu -q -k "int8 and $v" --timeout 100 2>&1 | tail -4; done
} 2>&1 | tee gpurun_out/r2_f64_i8.txt

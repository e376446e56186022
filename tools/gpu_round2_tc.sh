#!/bin/bash
# Round 2, tcgen05 SGEMM kernel variants written after round 1's GPU budget ran out (one B200, ~4 min):
#   make -C tools && gpurun --timeout 420 -- 'bash tools/gpu_round2_tc.sh'
# Every pipeline wait in these kernels is the guarded kind: a broken protocol traps after ~10 s (launch error), it does not hang the GPU.
#   TMM_TC_ATMEM=1        A operand through tensor memory (sgemm_tc_ts_kernel<false>)      224 -> 144 KB of smem traffic per k-block
#   TMM_TC_ATMEM=2        ... plus CTA pairs, cta_group::2 (sgemm_tc_ts_kernel<true>)       -> 88 KB per CTA and k-block
#   TMM_TC_TF32_STAGES=6  six-stage ring for the plain-TF32 mode (sgemm_tc_deep_kernel)     latency-bound with three stages
# What passes `check` for all four op pairs and is faster gets promoted to the default in gemm_f32_tc.cu (sgemm_tc_launch).
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
T=./build/tc_test
{
nvidia-smi -L | head -1
echo "== unit probes: tcgen05.st/ld round trip, tf32 MMA with A from smem / from TMEM, i8 MMA (one CTA, one MMA each) =="; timeout 60 ./build/tc_probe2
echo "== default kernel (reference point) =="; timeout 60 $T benchone N N 8192 8192 8192 0
echo "== sgemm, A operand through tensor memory (TMM_TC_ATMEM=1: 144 instead of 224 KB of shared-memory traffic per k-block) =="
for tt in "N N" "T N"; do TMM_TC_ATMEM=1 timeout 60 $T probe $tt 2>&1 | head -40; done   # single tile, single k-block: which A element reached which accumulator position (silent = all right)
for tt in "N N" "T N" "N T" "T T"; do TMM_TC_ATMEM=1 timeout 120 $T check $tt 2>&1 | grep -v " OK$" | tail -5; done
TMM_TC_ATMEM=1 timeout 120 $T precision 2>&1 | grep -E "precision|tmm fp32" | head -20
for tt in "N N" "T N" "N T" "T T"; do TMM_TC_ATMEM=1 timeout 60 $T benchone $tt 8192 8192 8192 0; done
TMM_TC_ATMEM=1 TMM_TC_SPLIT=trunc timeout 60 $T benchone N N 8192 8192 8192 0
echo "== sgemm, A through tensor memory + CTA pairs (TMM_TC_ATMEM=2: cta_group::2, 88 KB per CTA and k-block) =="
for tt in "N N" "T N"; do TMM_TC_ATMEM=2 timeout 60 $T probe $tt 2>&1 | head -40; done   # one pair tile: rows 0-127 only (rank 1 works on zero-filled rows)
for tt in "N N" "T N" "N T" "T T"; do TMM_TC_ATMEM=2 timeout 120 $T check $tt 2>&1 | grep -v " OK$" | tail -5; done
TMM_TC_ATMEM=2 timeout 120 $T precision 2>&1 | grep -E "precision|tmm fp32" | head -20
for tt in "N N" "T T"; do TMM_TC_ATMEM=2 timeout 60 $T benchone $tt 8192 8192 8192 0; done
TMM_TC_ATMEM=2 TMM_TC_SPLIT=trunc timeout 60 $T benchone N N 8192 8192 8192 0
echo "== plain TF32 mode (one MMA per product): 3 stages (measured 352 TF) vs six 32 KB stages (TMM_TC_TF32_STAGES=6) =="
for tt in "N N" "T T"; do TMM_TC_TF32_STAGES=6 timeout 120 $T check $tt 2>&1 | grep -E "tf32-mode|FAIL" | tail -3; done   # check ends with a TF32-mode case (expect ~1e-4)
TMM_TC_TF32_STAGES=6 timeout 60 $T benchone N N 8192 8192 8192 0   # the "tmm tf32" column; compare with the default run above
echo "== gated pytest for the variants (one process per variant: a trap poisons its CUDA context) =="
for v in one-cta cta-pair; do TMM_EXPERIMENTAL=1 timeout 150 python -m pytest tests/test_experimental_gpu.py -m gpu -q -k "tmem and $v" --timeout 100 2>&1 | tail -4; done
} 2>&1 | tee gpurun_out/r2_tc_variants.txt

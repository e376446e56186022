#!/bin/bash
# First GPU call of round 2 (one B200, ~6 min): validate everything that was written after round 1's GPU budget ran out.
#   make -C tools && gpurun --timeout 500 -- 'bash tools/gpu_round2.sh'
# 1. gated experimental tests (tcgen05 CGEMM embedding, BF16 entry points, device-pointer operands)            -> promote TMM_C32_MATH=tc to the default if green
# 2. tc_test cgemm: all nine op pairs vs cuBLAS CGEMM + timing      -> CGEMM number for DESIGN 3.4
# 3. SGEMM split variants: precision + timing, default vs TMM_TC_SPLIT=trunc   (the new tcgen05 kernel variants: tools/gpu_round2_tc.sh)
# 4. the regular suite + bench line (regression check)
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
T=./build/tc_test
{
nvidia-smi -L | head -1
echo "== experimental pytest =="; TMM_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_experimental_gpu.py -m gpu -q --timeout 120 -k "not tmem and not int8" 2>&1 | tail -25   # the kernel variants have their own scripts (a trap would poison this process)
echo "== native bf16 kind::f16 (TN) =="; TMM_BF16_NATIVE=1 TMM_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_experimental_gpu.py -m gpu -q -k bf16 --timeout 100 2>&1 | tail -6
echo "== tc_test cgemm =="; timeout 240 $T cgemm 2>&1 | grep -v " OK$" | tail -40
echo "== sgemm split: round-to-nearest (default) =="; timeout 120 $T precision 2>&1 | grep -E "precision|tmm fp32|cuBLAS fp32 pedantic" | head -20
timeout 60 $T benchone N N 8192 8192 8192 0
echo "== sgemm split: hi = raw bits (TMM_TC_SPLIT=trunc) =="; TMM_TC_SPLIT=trunc timeout 120 $T precision 2>&1 | grep -E "precision|tmm fp32" | head -20
TMM_TC_SPLIT=trunc timeout 60 $T benchone N N 8192 8192 8192 0
for tt in "N N" "T T"; do TMM_TC_SPLIT=trunc timeout 120 $T check $tt 2>&1 | grep -v " OK$" | tail -5; done
echo "== published experiment (README figure: dgemm square, alpha=beta=1), both arms =="; timeout 400 python tools/sweep_published.py --reps 2 2>&1 | tail -12
echo "== beta = 1 at 10000^3: default stripes vs one C stripe per k-chunk =="
timeout 60 python tools/e2e.py --beta 1 --reps 4 2>&1 | tail -1
TMM_PLAN_CSTRIPES=chunks timeout 60 python tools/e2e.py --beta 1 --reps 4 2>&1 | tail -1
echo "== mid-size products: n1 bound on / off =="
for n in 4000 6000 8000; do timeout 60 python tools/e2e.py --m $n --n $n --k $n --reps 5 2>&1 | tail -1; TMM_PLAN_D2H_BOUND=0 timeout 60 python tools/e2e.py --m $n --n $n --k $n --reps 5 2>&1 | tail -1; done
echo "== cublasXt comparator (the reference's headline comparison), tuned block 4000 and default =="
for n in 4000 10000 16000; do timeout 120 ./build/cublasxt-multiply -m $n -n $n -k $n -r 2 --beta 1 --block 4000 2>&1 | grep -E "Avg Time|Throughput"; timeout 60 bin/multiply -m $n -n $n -k $n -r 2 --beta 1 2>&1 | grep -E "Avg Time|Throughput" | head -2; done
echo "== SGEMM host-to-host: planner with the float rate (default) vs the round-1 FP64 rate =="
for n in 10000 16000; do timeout 60 python tools/e2e.py --dtype s --m $n --n $n --k $n --reps 4 2>&1 | tail -1; TMM_PLAN_F32_FLOPS=35e12 TMM_PLAN_D2H_BOUND=0 timeout 60 python tools/e2e.py --dtype s --m $n --n $n --k $n --reps 4 2>&1 | tail -1; done
echo "== the reference's DEFAULT call mode (pin_host_buffers=true on pageable memory): 1 vs 4 vs 8 registration threads, and the reference =="
for t in 1 4 8; do TMM_PIN_THREADS=$t timeout 120 python tools/pin_study.py 10000 2>&1 | tail -2 | head -1; done; timeout 120 python tools/pin_study.py 10000 2>&1 | tail -1
echo "== compute-sanitizer memcheck on the CI shapes (SURVEY 5.2) =="; timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gemm_gpu.py -m gpu -q -k "ci_and_ctest or degenerate" 2>&1 | tail -6
echo "== regular suite =="; timeout 600 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8
echo "== bench =="; timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r2_first.txt

#!/bin/bash
# Round 2: the beta != 0 call (the reference's published experiment: alpha = beta = 1) with the caller's C added at the END of each block's
# accumulation (its upload no longer gates the first GEMM) against round 1's C-first order.  (one B200)
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
echo "##### pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
echo "##### beta = 1, 10000^3: C deferred (default) vs C first (TMM_PLAN_DEFER_C=0)"
timeout 90 python tools/e2e.py --beta 1 --reps 5 2>&1 | tail -1
TMM_PLAN_DEFER_C=0 timeout 90 python tools/e2e.py --beta 1 --reps 5 2>&1 | tail -1
echo "##### beta = 1, other sizes and ops: deferred vs first"
for n in 4000 8000 16000; do timeout 120 python tools/e2e.py --beta 1 --reps 4 --m $n --n $n --k $n 2>&1 | tail -1; TMM_PLAN_DEFER_C=0 timeout 120 python tools/e2e.py --beta 1 --reps 4 --m $n --n $n --k $n 2>&1 | tail -1; done
timeout 90 python tools/e2e.py --beta 1 --reps 4 --tt TN --dtype z --m 6000 --n 6000 --k 6000 2>&1 | tail -1
TMM_PLAN_DEFER_C=0 timeout 90 python tools/e2e.py --beta 1 --reps 4 --tt TN --dtype z --m 6000 --n 6000 --k 6000 2>&1 | tail -1
echo "##### beta = 0, 10000^3 (unchanged path)"; timeout 90 python tools/e2e.py --reps 5 2>&1 | tail -1
echo "##### opt-in int8 mode after the slicing store fix: device-resident and host-to-host"
TMM_F64_MATH=i8:7 timeout 90 ./build/devtest benchone N N 10000 10000 10000 0
TMM_F64_MATH=i8:7 timeout 90 python tools/e2e.py --reps 5 2>&1 | tail -1
TMM_F64_MATH=i8:7 timeout 90 python tools/e2e.py --reps 5 --beta 1 2>&1 | tail -1
echo "##### trace, beta = 1"
TMM_TRACE=1 timeout 90 python tools/e2e.py --beta 1 --reps 2 2>&1 | grep "trace\]" | tail -70 | sed 's/\[tmm trace\] //' | grep -v "gemm1(1\|gemm1(2\|gemm1(4"
} 2>&1 | tee gpurun_out/r2_beta1.txt

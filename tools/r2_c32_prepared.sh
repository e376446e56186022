#!/bin/bash
# Round 2: complex<float> host to host with operands prepared once in context buffers (no stream-ordered allocation on the launch path):
# float-type parity, timings in fresh processes and in the process that also runs the reference (where the pool-based path went from 30 to 89 ms).  (one B200)
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
echo "##### pytest -m gpu: float / complex<float> tests"
timeout 300 python -m pytest tests -m gpu -x -q --timeout 200 -k "float or cgemm or sgemm or device_gemm_boundary or golden or own_test_multiply or streaming or sweep" 2>&1 | tail -3
for n in 4000 8000 10000; do
  echo "##### host to host cgemm $n^3"
  timeout 120 python tools/e2e.py --dtype c --m $n --n $n --k $n --reps 5 2>&1 | tail -1
done
timeout 120 python tools/e2e.py --dtype c --tt CT --m 8000 --n 8000 --k 8000 --reps 4 --beta 1 2>&1 | tail -1
echo "##### complex<float> against the unmodified reference, both in one process (beta = 0, then the published alpha = beta = 1)"
timeout 200 python tools/sweep_published.py --dtype c --sizes 4000,8000,10000 --beta 0 --reps 3 2>&1 | tail -4
timeout 200 python tools/sweep_published.py --dtype c --sizes 6000,10000 --beta 1 --reps 3 2>&1 | tail -3
echo "##### trace, cgemm 8000^3 (last call of 3)"
TMM_TRACE=1 timeout 120 python tools/e2e.py --dtype c --m 8000 --n 8000 --k 8000 --reps 3 --fill const > gpurun_out/r2_cgemm_trace_prepared.txt 2>&1; tail -1 gpurun_out/r2_cgemm_trace_prepared.txt
} 2>&1 | tee gpurun_out/r2_c32_prepared.txt

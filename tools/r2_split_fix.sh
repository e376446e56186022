#!/bin/bash
# Round 2: the TF32 split of the FP32-accurate tcgen05 kernel takes a five-instruction path for chunks of plain values (the special cases cost the
# mode 155 -> 96 TF when they were applied to every element).  Float-type parity, the kernel alone, complex<float> / float host to host with the
# dynamic tile feed and with the fixed stride, and the four types against the unmodified reference.  (one B200)
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
echo "##### pytest -m gpu: float / complex<float> tests"
timeout 300 python -m pytest tests -m gpu -x -q --timeout 200 -k "float or cgemm or bf16 or sgemm or device_gemm_boundary or golden or own_test_multiply" 2>&1 | tail -3
echo "##### device-resident tcgen05 kernels 8192^3"
timeout 60 ./build/tc_test benchone N N 8192 8192 8192 0 2>&1 | tail -1
timeout 60 ./build/tc_test benchone T N 8192 8192 8192 0 2>&1 | tail -1
for n in 4000 8000 10000; do
  echo "##### host to host cgemm $n^3: dynamic tile feed (default) | static stride"
  timeout 120 python tools/e2e.py --dtype c --m $n --n $n --k $n --reps 5 --fill const 2>&1 | tail -1
  TMM_TC_SCHED=static timeout 120 python tools/e2e.py --dtype c --m $n --n $n --k $n --reps 5 --fill const 2>&1 | tail -1
done
echo "##### the four types against the unmodified reference (beta = 0)"
for t in s c; do timeout 200 python tools/sweep_published.py --dtype $t --sizes 4000,8000,10000 --beta 0 --reps 3 2>&1 | tail -5; done
echo "##### trace, cgemm 8000^3 (last call of 3)"
TMM_TRACE=1 timeout 120 python tools/e2e.py --dtype c --m 8000 --n 8000 --k 8000 --reps 3 --fill const > gpurun_out/r2_cgemm_trace_split_fix.txt 2>&1; tail -1 gpurun_out/r2_cgemm_trace_split_fix.txt
} 2>&1 | tee gpurun_out/r2_split_fix.txt

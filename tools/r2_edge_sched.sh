#!/bin/bash
# Round 2: (1) the FP64 kernels skip DMMA tiles that lie past the edge of C, (2) the tcgen05 float kernel draws its tiles from a per-launch counter
# (TMM_TC_SCHED=static: the fixed stride it replaces), (3) the library's own scratch pool.  Parity first, then timings.  (one B200)
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
nvidia-smi -L
echo "##### pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
echo "##### device-resident kernels: dgemm 10000^3 (vs cuBLAS), zgemm via kbench"
timeout 90 ./build/devtest benchone N N 10000 10000 10000 0 2>&1 | tail -3
echo "##### host to host dgemm 10000^3"; timeout 90 python tools/e2e.py --reps 8 2>&1 | tail -1
echo "##### host to host zgemm 10000^3"; timeout 120 python tools/e2e.py --dtype z --reps 4 --fill const 2>&1 | tail -1
for n in 4000 8000 10000; do
  echo "##### host to host cgemm $n^3: dynamic tile feed (default) | static stride"
  timeout 120 python tools/e2e.py --dtype c --m $n --n $n --k $n --reps 5 --fill const 2>&1 | tail -1
  TMM_TC_SCHED=static timeout 120 python tools/e2e.py --dtype c --m $n --n $n --k $n --reps 5 --fill const 2>&1 | tail -1
done
echo "##### device-resident tcgen05 kernels 8192^3: dynamic | static"
timeout 60 ./build/tc_test benchone N N 8192 8192 8192 0 2>&1 | tail -2
TMM_TC_SCHED=static timeout 60 ./build/tc_test benchone N N 8192 8192 8192 0 2>&1 | tail -2
echo "##### trace, cgemm 8000^3 (last call of 3)"
TMM_TRACE=1 timeout 120 python tools/e2e.py --dtype c --m 8000 --n 8000 --k 8000 --reps 3 --fill const > gpurun_out/r2_cgemm_trace_dynamic.txt 2>&1; tail -1 gpurun_out/r2_cgemm_trace_dynamic.txt
echo "##### bench.py (ours)"; timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1
} 2>&1 | tee gpurun_out/r2_edge_sched.txt

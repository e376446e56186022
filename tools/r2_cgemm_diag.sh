#!/bin/bash
# Round 2: why the host-to-host complex<float> call trails the reference at n <= 8000 (profiles/r2_types_sweep.txt: 0.88x / 0.91x) although the
# tcgen05 CGEMM is 1.8x cuBLAS CGEMM device-resident.  Traces the call with the library's own scratch pool (default) and with the device's default
# pool (TMM_POOL_KEEP=0: free blocks go back to the driver at every synchronisation).  (one B200)
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
for n in 4000 8000; do
  echo "##### cgemm $n^3, library pool (default)"; timeout 120 python tools/e2e.py --dtype c --m $n --n $n --k $n --reps 5 --fill const 2>&1 | tail -6
  echo "##### cgemm $n^3, TMM_POOL_KEEP=0";        TMM_POOL_KEEP=0 timeout 120 python tools/e2e.py --dtype c --m $n --n $n --k $n --reps 5 --fill const 2>&1 | tail -6
done
echo "##### sgemm 8000^3 (same scheduler, no operand preparation)"; timeout 120 python tools/e2e.py --dtype s --m 8000 --n 8000 --k 8000 --reps 4 --fill const 2>&1 | tail -1
echo "##### trace, cgemm 8000^3, library pool (last call of 3)"
TMM_TRACE=1 timeout 120 python tools/e2e.py --dtype c --m 8000 --n 8000 --k 8000 --reps 3 --fill const > gpurun_out/r2_cgemm_trace_keep.txt 2>&1
grep -c "trace\]" gpurun_out/r2_cgemm_trace_keep.txt
echo "##### trace, cgemm 8000^3, TMM_POOL_KEEP=0 (last call of 3)"
TMM_POOL_KEEP=0 TMM_TRACE=1 timeout 120 python tools/e2e.py --dtype c --m 8000 --n 8000 --k 8000 --reps 3 --fill const > gpurun_out/r2_cgemm_trace_nokeep.txt 2>&1
grep -c "trace\]" gpurun_out/r2_cgemm_trace_nokeep.txt
} 2>&1 | tee gpurun_out/r2_cgemm_diag.txt

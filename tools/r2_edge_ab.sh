#!/bin/bash
# Round 2: do the FP64 kernels gain from skipping the DMMA tiles past the edge of C?  Same binary, the library at HEAD against a variant whose
# gemm_f64.cu is the one before the change (build/variants/f64old), on shapes with 1.6 % (10000^3) and ~10 % (m = 1040) of edge padding.  (one B200)
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
for shape in "1040 20000 10000" "10000 10000 10000"; do
  echo "##### $shape: HEAD (edge tiles skipped) | before"
  timeout 40 ./build/devtest benchone N N $shape 0 2>&1 | tail -1
  LD_LIBRARY_PATH=build/variants/f64old:$LD_LIBRARY_PATH timeout 40 ./build/devtest benchone N N $shape 0 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/r2_edge_ab.txt

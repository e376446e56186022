#!/bin/bash
# Round 2: which change to the tcgen05 float kernel costs the 3xTF32 mode its pace (155 -> 96 TF at 8192^3 after the tile ring went in)?
# Variant libraries under build/variants/ (built by hand from the kernel file with one thing changed each), same tc_test binary.  (one B200)
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
for v in old r40 ring0; do
  [ -f build/variants/$v/libtiledmm_b200.so ] || continue
  echo "##### variant $v"
  LD_LIBRARY_PATH=build/variants/$v:$LD_LIBRARY_PATH timeout 60 ./build/tc_test benchone N N 8192 8192 8192 0 2>&1 | tail -1
done
echo "##### HEAD library"
timeout 60 ./build/tc_test benchone N N 8192 8192 8192 0 2>&1 | tail -1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv
} 2>&1 | tee gpurun_out/r2_bisect_tc.txt

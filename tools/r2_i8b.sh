#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
for c in 1 0; do for g in 2 4 8 16; do echo "clusters $c group $g"; TMM_DEBUG=1 TMM_I8_GROUP=$g TMM_I8_CLUSTER=$c TMM_F64_MATH=i8:7 timeout 90 ./build/devtest benchone N N 10000 10000 10000 0 2>&1 | grep -E "bench|clusters of"; done; done
} 2>&1 | tee gpurun_out/r2_i8b.txt

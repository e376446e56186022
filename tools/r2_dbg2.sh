#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
TMM_DEBUG=1 timeout 300 bin/multiply -m 20000 -n 20000 -k 20000 --scaling 2 -r 2 --random 1 2>&1 | grep -v "push #" | tail -60
TMM_TRACE=1 timeout 300 bin/multiply -m 20000 -n 20000 -k 20000 --scaling 2 -r 2 --random 1 2>&1 | grep -E "trace\] (host|h2dAB\(0|gemm1\(0,[0-9]+,256|d2hC|h2dB|gemm2)|SCALING|call" | tail -80
} 2>&1 | tee gpurun_out/r2_dbg2.txt

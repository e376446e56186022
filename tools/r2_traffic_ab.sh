#!/bin/bash
# Round 2: is the DRAM traffic of the DGEMM launch (21.3 GB at HEAD, 13.2 GB in the capture of the build before) a property of the build or of the run?
# The same four counters for the library at HEAD and for the variant with the previous gemm_f64.cu, back to back.  (one B200)
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active
{
echo "##### HEAD"
timeout 25 ncu --metrics $M --clock-control none -k regex:dgemm_kernel -s 2 -c 1 ./build/devtest benchone N N 10000 10000 10000 0 2>&1 | grep -E "dram__|lts__|gpu__time|dmma"
echo "##### previous gemm_f64.cu"
LD_LIBRARY_PATH=build/variants/f64old:$LD_LIBRARY_PATH timeout 25 ncu --metrics $M --clock-control none -k regex:dgemm_kernel -s 2 -c 1 ./build/devtest benchone N N 10000 10000 10000 0 2>&1 | grep -E "dram__|lts__|gpu__time|dmma"
} 2>&1 | tee gpurun_out/r2_traffic_ab.txt

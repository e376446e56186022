#!/bin/bash
# Round 2, final single-GPU run at HEAD: full GPU suite, smoke, both bench arms; then ncu captures of the two hot kernels of this build
# (DGEMM for bench.py's roofline.traffic, the FP32-accurate tcgen05 kernel after the split fix).
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
nvidia-smi -L
echo "##### pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
echo "##### smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
echo "##### bench.py --impl reference"; timeout 300 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1
echo "##### bench.py (ours)"; timeout 300 python bench.py --steps 20 --warmup 5 2>&1 | tail -1
echo "##### the tcgen05 float kernel alone (every change to the split stage or the kernel is timed: DESIGN 8, item 2)"; timeout 60 ./build/tc_test benchone N N 8192 8192 8192 0 2>&1 | tail -1
NCU="ncu --set full --clock-control none --import-source on"
echo "##### ncu --set full: dgemm_kernel 10000^3, sgemm_tc_kernel (3xTF32) 8192^3"
timeout 120 $NCU -k regex:dgemm_kernel -s 2 -c 1 -o gpurun_out/r2_prof_dgemm_head -f ./build/devtest benchone N N 10000 10000 10000 0 2>&1 | tail -2
timeout 120 $NCU -k regex:sgemm_tc_kernel -s 2 -c 1 -o gpurun_out/r2_prof_sgemm_fp32_head -f ./build/tc_test benchone N N 8192 8192 8192 0 2>&1 | tail -2
ls -la gpurun_out/*head*.ncu-rep
} 2>&1 | tee gpurun_out/r2_final3.txt

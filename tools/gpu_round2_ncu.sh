#!/bin/bash
# ncu evidence for round 2 (one B200, ~5 min): full captures of the tcgen05 SGEMM kernel (FP32-accurate and plain TF32) and of the
# DGEMM kernel, plus the launch list of bench.py.  Read the .ncu-rep files back here with `ncu -i ... --page raw --csv`.
#   make -C tools && gpurun --timeout 420 -- 'bash tools/gpu_round2_ncu.sh'
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
NCU="ncu --set full --clock-control none --import-source on"
{
timeout 150 $NCU -k regex:sgemm_tc_kernel -s 2 -c 1 -o gpurun_out/r2_prof_sgemm_fp32 -f ./build/tc_test benchone N N 8192 8192 8192 0 2>&1 | tail -3
TMM_F32_MATH=tf32 timeout 150 $NCU -k regex:sgemm_tc_kernel -s 2 -c 1 -o gpurun_out/r2_prof_sgemm_tf32 -f ./build/tc_test benchone N N 8192 8192 8192 0 2>&1 | tail -3
timeout 200 $NCU -k regex:dgemm_kernel -s 4 -c 1 -o gpurun_out/r2_prof_dgemm -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -3
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
} 2>&1 | tee gpurun_out/r2_ncu.txt

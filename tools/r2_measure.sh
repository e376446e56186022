#!/bin/bash
# Round 2, one B200: schedule sweep, the rest of the miniapp (device-resident C, beta = 1, pin_host_buffers = true), the reference's published
# experiment on both arms with the cublasXt comparator, ncu captures.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
nvidia-smi -L
echo "##### schedule knobs, dgemm 10000^3 beta = 0"; timeout 400 python tools/sweep_plan.py --reps 6 2>&1 | tail -30
} 2>&1 | tee gpurun_out/r2_sweep_plan.txt
{
echo "##### published experiment (README.md:20-25): dgemm square, alpha = beta = 1, both arms"; timeout 500 python tools/sweep_published.py --reps 2 --sizes 4000,8000,10000,12000,16000,20000,24000,28000,32000 2>&1 | tail -14
echo "##### the miniapp's second variant: C stays on the device (copy_c_back = false), beta = 0"; timeout 300 python tools/sweep_published.py --reps 2 --beta 0 --copy-c-back 0 --sizes 4000,10000,16000 2>&1 | tail -6
echo "##### beta = 0, copy back (the headline shape at other sizes)"; timeout 300 python tools/sweep_published.py --reps 2 --beta 0 --sizes 4000,8000,10000,16000 2>&1 | tail -7
echo "##### cublasXt comparator (examples/cublasXt-multiply.cpp), tuned block 4000, alpha = beta = 1"
for n in 4000 10000 16000 32000; do timeout 200 ./build/cublasxt-multiply -m $n -n $n -k $n -r 2 --beta 1 --block 4000 2>&1 | grep -E "Avg Time|Throughput" | tr '\n' ' '; echo " (n = $n)"; done
echo "##### pin_host_buffers = true on pageable memory (the API default)"; timeout 300 python tools/pin_study.py 10000 2>&1 | grep -v "run"
} 2>&1 | tee gpurun_out/r2_published_sweep.txt
NCU="ncu --set full --clock-control none --import-source on"
{
timeout 150 $NCU -k regex:sgemm_tc_kernel -s 2 -c 1 -o gpurun_out/r2_prof_sgemm_fp32 -f ./build/tc_test benchone N N 8192 8192 8192 0 2>&1 | tail -3
TMM_F32_MATH=tf32 timeout 150 $NCU -k regex:sgemm_tc_kernel -s 2 -c 1 -o gpurun_out/r2_prof_sgemm_tf32 -f ./build/tc_test benchone N N 8192 8192 8192 0 2>&1 | tail -3
timeout 200 $NCU -k regex:dgemm_kernel -s 2 -c 1 -o gpurun_out/r2_prof_dgemm -f ./build/devtest benchone N N 10000 10000 10000 0 2>&1 | tail -3
TMM_F64_MATH=i8:7 timeout 200 $NCU -k regex:dgemm_i8_kernel -s 1 -c 1 -o gpurun_out/r2_prof_dgemm_i8 -f ./build/devtest benchone N N 10000 10000 10000 0 2>&1 | tail -3
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
} 2>&1 | tee gpurun_out/r2_ncu.txt

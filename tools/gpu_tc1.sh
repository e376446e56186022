#!/bin/bash
# GPU-box pass for the tcgen05 SGEMM: layout probes, descriptor sweep, correctness, timing, one ncu capture.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
T=./build/tc_test
{
nvidia-smi -L | head -1
fail=0
for tt in "T N" "N N" "N T" "T T"; do
  echo "== probe $tt =="; timeout 60 $T probe $tt; rc=$?; echo "rc=$rc"; [ $rc -ne 0 ] && fail=1
done
if [ $fail -ne 0 ]; then
  echo "== MN descriptor sweep (probe N T: both operands MN-major) =="
  for lay in "1 4" "2 3"; do set -- $lay
    for ls in "4096 512" "512 4096" "4096 1024" "1024 4096" "4096 4096" "512 512"; do set -- $lay $ls
      echo "-- layout=$1 tma_swizzle=$2 lbo=$3 sbo=$4"
      TMM_TC_MN_LAYOUT=$1 TMM_TC_MN_SWIZZLE=$2 TMM_TC_MN_LBO=$3 TMM_TC_MN_SBO=$4 timeout 60 $T probe N T 2>&1 | grep -E "probe|rc=|error"
    done
  done
fi
for tt in "T N" "N N" "N T" "T T"; do
  echo "== check $tt =="; timeout 180 $T check $tt; echo "rc=$?"
done
echo "== bench =="; timeout 240 $T bench; echo "rc=$?"
echo "== host =="; timeout 240 $T host; echo "rc=$?"
} 2>&1 | tee gpurun_out/tc1.txt
if grep -q "check N N: 0 failing" gpurun_out/tc1.txt; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:sgemm_tc_kernel -s 1 -c 1 -o gpurun_out/prof_sgemm_tc $T benchone N N 8192 8192 8192 0 > gpurun_out/ncu_sgemm_tc.log 2>&1
  tail -3 gpurun_out/ncu_sgemm_tc.log
fi
ls -la gpurun_out

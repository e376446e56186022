#!/bin/bash
# GPU-box pass 2 for the tcgen05 SGEMM: windowed promotion - correctness, accuracy vs window, speed vs window.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
T=./build/tc_test
{
nvidia-smi -L | head -1; nproc
for tt in "N N" "T N" "N T" "T T"; do
  echo "== check $tt =="; timeout 180 $T check $tt 2>&1 | grep -v " OK$"; echo "rc=$?"
done
for w in 1 2 4 8 100000; do
  echo "== precision window=$w =="; TMM_TC_WINDOW=$w timeout 300 $T precision
done
for w in 1 2 4 8 100000; do
  echo "== bench window=$w =="; TMM_TC_WINDOW=$w timeout 120 $T benchone N N 8192 8192 8192 0; TMM_TC_WINDOW=$w timeout 120 $T benchone N N 10000 4800 512 1
done
echo "== bench =="; timeout 240 $T bench
echo "== host =="; timeout 240 $T host
} 2>&1 | tee gpurun_out/tc2.txt

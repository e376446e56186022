#!/bin/bash
# GPU-box pass 3: tcgen05 SGEMM with 4 TMEM window accumulators; full GPU parity suite; smoke.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
T=./build/tc_test
{
nvidia-smi -L | head -1
for tt in "N N" "T N" "N T" "T T"; do
  echo "== check $tt =="; timeout 180 $T check $tt 2>&1 | grep -v " OK$"
done
echo "== bench =="; timeout 240 $T bench
echo "== host =="; timeout 240 $T host
echo "== pytest gpu =="; timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -15
echo "== smoke =="; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5
} 2>&1 | tee gpurun_out/tc3.txt

#!/bin/bash
# Round 2 multi-GPU run: tests/test_multi_gpu.py, the bench line at N = G (and smaller counts), the strong-scaling miniapp.
#   gpurun --gpus G --timeout T -- 'bash tools/r2_multi.sh G [c5]'
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
G=${1:-2}
MODE=${2:-quick}
{
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1
echo "##### pytest tests/test_multi_gpu.py -m gpu"
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 300 2>&1 | tail -6
for n in $G $((G/2)); do
  [ "$n" -lt 2 ] && continue
  [ "$n" -lt 4 ] && [ "$G" -ge 8 ] && continue
  echo "##### bench.py --gpus $n"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 5 --warmup 3 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -3
done
if [ "$MODE" = "bench" ]; then
  echo "##### bench.py --impl reference --gpus $G (rank 0 runs the same global problem on one GPU)"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29650 bench.py --impl reference --gpus $G --steps 3 --warmup 1 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -2
else
echo "##### strong scaling through the drop-in C++ path (one process, a host thread per GPU): dgemm 20000^3"
timeout 300 bin/multiply -m 20000 -n 20000 -k 20000 --scaling $G,$G,1 --random 1 2>&1 | grep -E "SCALING|SPEEDUP|host buffers|rror"
fi
if [ "$MODE" = "c5" ]; then
  echo "##### BASELINE configs[4]: dgemm 100000^3 (240 GB, out of core on one GPU)"
  TMM_DIST_TIMEOUT_S=120 timeout 900 bin/multiply -m 100000 -n 100000 -k 100000 --scaling $G,$G,1 --random 1 2>&1 | grep -E "SCALING|SPEEDUP|host buffers|rror"
  echo "##### BASELINE configs[3]: zgemm 20000 x 20000 x 500000 at reduced k = 100000 ($G GPUs and 1), then full k on $G GPUs"
  TMM_DIST_TIMEOUT_S=120 timeout 300 bin/multiply --type z -m 20000 -n 20000 -k 100000 --scaling $G,1 --random 1 2>&1 | grep -E "SCALING|SPEEDUP|host buffers|rror"
  TMM_DIST_TIMEOUT_S=120 timeout 400 bin/multiply --type z -m 20000 -n 20000 -k 500000 --scaling $G --random 1 2>&1 | grep -E "SCALING|SPEEDUP|host buffers|rror"
fi
} 2>&1 | tee gpurun_out/r2_multi_${G}gpu_${MODE}.txt

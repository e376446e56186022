#!/bin/bash
# Round 2, 8 GPUs: upload shares that follow the measured link rates (default) against equal shares (TMM_DIST_BALANCE=0).
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
echo "##### shares chosen (one grid call at the bench's N = 8 block shape, debug lines of every rank)"
TMM_DEBUG=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29701 tools/e2e.py --fill const --m 10000 --n 5000 --k 20000 --reps 3 2>&1 | grep -E "upload shares|host link with|E2E|run " | sort -u | head -40
echo "##### bench.py --gpus 8, balanced shares (default)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 8 --steps 8 --warmup 3 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -2
echo "##### bench.py --gpus 8, equal shares (TMM_DIST_BALANCE=0)"
TMM_DIST_BALANCE=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29703 bench.py --gpus 8 --steps 8 --warmup 3 2>&1 | grep -E '^\{|Error|error|Traceback' | tail -2
} 2>&1 | tee gpurun_out/r2_balance8.txt

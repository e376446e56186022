#!/bin/bash
# Round 2, 8-GPU diagnostic: what do 8 concurrent host links deliver on this node, and where does the grid call wait?
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
{
nvidia-smi -L | wc -l; nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core" ; nvidia-smi topo -m | head -12
echo "##### probe_multi"; timeout 120 ./build/probe_multi 256
echo "##### torchrun 8 ranks, rank block 10000^3 (weak; the round-1 bench workload), all ranks traced"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/e2e.py --fill const --reps 4 --trace-dir gpurun_out/diag8_trace 2>&1 | grep -v "^\[tmm trace\]" | tail -12
echo "##### torchrun 4 ranks"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 tools/e2e.py --fill const --reps 4 --trace-dir gpurun_out/diag4_trace 2>&1 | grep -v "^\[tmm trace\]" | tail -8
echo "##### 8 ranks, NCCL data plane"
TMM_DIST_NCCL=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/e2e.py --fill const --reps 4 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2_diag8.txt

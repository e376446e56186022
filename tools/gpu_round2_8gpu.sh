#!/bin/bash
# 8-GPU call of round 2 (charged 8x: keep it short): the 2x4 grid through torchrun, then the single-process drop-in path.
#   gpurun --gpus 8 --timeout 150 -- 'bash tools/gpu_round2_8gpu.sh'
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
{
nvidia-smi -L | wc -l
TMM_DIST_TIMEOUT_S=30 timeout -k 3 70 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus 8 --steps 4 --warmup 3 2>&1 | tail -2
timeout 40 bin/multiply -m 20000 -n 40000 -k 10000 -r 2 --gpus 8 2>&1 | grep -E "Avg Time|Throughput|last call" | head -3
} 2>&1 | tee gpurun_out/r2_8gpu.txt

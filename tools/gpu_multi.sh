#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): grid tests + N-rank bench.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi -L | tee gpurun_out/box_multi.txt; nproc >> gpurun_out/box_multi.txt; free -g | head -2 >> gpurun_out/box_multi.txt
nvidia-smi topo -m 2>&1 | head -14 >> gpurun_out/box_multi.txt
echo "== pytest multi gpu =="; timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q --timeout 180 2>&1 | tail -25 | tee gpurun_out/pytest_multi.txt
echo "== bench N=$N =="; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | tail -4 | tee gpurun_out/bench_n$N.json

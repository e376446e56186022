#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): grid tests + N-rank bench + single-process N-device call.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${N:-2}
{
nvidia-smi -L; nproc; free -g | head -2; nvidia-smi topo -m 2>&1 | head -14
} > gpurun_out/box_multi_n$N.txt 2>&1
echo "== pytest multi gpu =="; timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -x -q --timeout 180 2>&1 | tail -25 | tee gpurun_out/pytest_multi_n$N.txt
echo "== bench N=$N =="; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n$N.json
echo "== single process, $N devices, strong scaling =="; timeout 300 python tools/e2e.py --devices $N --m ${SZ:-20000} --n ${SZ:-20000} --k ${SZ:-20000} --reps 4 2>&1 | grep -E "run|E2E" | tee gpurun_out/e2e_single_n$N.txt

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${N:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/e2e.py --reps 4 --trace 2>&1 | grep -E "trace|run|E2E" | tail -100 | tee gpurun_out/trace_n$N.txt

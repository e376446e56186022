#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export TMM_DEBUG=1 NCCL_DEBUG=WARN
echo "== single process =="; timeout 150 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s -k single_process 2>&1 | tail -60 | tee gpurun_out/dbg_single.txt
echo "== per process =="; timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q -s -k one_process_per 2>&1 | tail -60 | tee gpurun_out/dbg_perproc.txt

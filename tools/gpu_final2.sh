#!/bin/bash
# 2-GPU evidence: the multi-GPU tests (single-process set_devices path and one-process-per-GPU torchrun path) and the N=2 bench line.
mkdir -p gpurun_out
OUT=gpurun_out/r1_final_2gpu.txt
{ nvidia-smi -L; nvidia-smi topo -m 2>/dev/null | head -6; date; } > $OUT
echo "== pytest tests/test_multi_gpu.py -m gpu ==" >> $OUT
( timeout 80 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 70 2>&1 | tail -15 ) >> $OUT
echo "== torchrun bench.py --gpus 2 ==" >> $OUT
( timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | tail -3 ) >> $OUT
date >> $OUT
tail -40 $OUT

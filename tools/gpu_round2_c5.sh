#!/bin/bash
# BASELINE configs[4] / [3] through the drop-in C++ path (one process, host thread per GPU, ONE copy of A, B, C in pinned host memory).
#   gpurun --gpus 8 --timeout 400 -- 'bash tools/gpu_round2_c5.sh 8'      (charged 8x: ~100 s of box time)
#   gpurun --timeout 400 -- 'bash tools/gpu_round2_c5.sh 1'               (the 1-GPU point of the strong-scaling pair: ~3.5 min)
# 100000^3 needs 240 GB of pinned host memory; the script falls back to 60000^3 (86 GB) when MemAvailable is short.
cd "${GRAFT_REPO_ROOT:-.}" || exit 1
mkdir -p gpurun_out
G=${1:-1}
AVAIL_GB=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
N=100000; [ "$AVAIL_GB" -lt 300 ] && N=60000
{
echo "gpus $G, MemAvailable ${AVAIL_GB} GB, n = $N"; nvidia-smi -L | wc -l
timeout 360 bin/multiply -m $N -n $N -k $N -r 1 --gpus $G --random 1 --variants back --warmup 0 2>&1 | grep -E "Avg Time|Throughput|last call|error|ERROR"
} 2>&1 | tee gpurun_out/r2_c5_${G}gpu.txt

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
python tools/nvlink_probe.py
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,P2P,SHM,NET timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/nvlink_probe.py 2>&1 | grep -E "all_gather| via |WARN" | sort | uniq -c | sort -rn | head -30
cat /proc/self/status | grep -i cap; ls -la /dev/shm | head -5; df -h /dev/shm | tail -1
} 2>&1 | tee gpurun_out/nvlink_probe.txt

#!/bin/bash
cd $GRAFT_REPO_ROOT
D=./build/devtest
echo "== kernel RMW cost =="
for k in 256 512 1024 2048 4096; do $D benchone N N 10000 5504 $k 1; $D benchone N N 10000 5504 $k 0; done
for i in 1 2 3; do
echo "== default rep $i =="; $D hostone N N 10000 10000 10000 0 1 2 6 2>&1 | grep -E "HOST|run"
done
echo "== TRACE default =="; TMM_TRACE=1 $D hostone N N 10000 10000 10000 0 1 2 2 2>&1 | grep -E "trace|run" | tail -56

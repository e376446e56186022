#!/bin/bash
cd $GRAFT_REPO_ROOT
D=./build/devtest
echo "== kernel small-K =="; 
$D benchone N N 10000 4800 256 1; $D benchone N N 10000 4800 512 1; $D benchone N N 10000 4800 1024 1; $D benchone N N 10000 4800 2048 1; $D benchone N N 10000 4736 256 1; $D benchone N N 10000 10000 256 1; $D benchone N N 10000 10000 512 1;  $D benchone N N 10000 512 10000 0; $D benchone N N 10000 192 10000 0; $D benchone N N 10000 256 10000 0
for mg in 1.3 1.5 1.8 2.5; do for kc0 in 256 512; do
echo "== margin $mg kc0 $kc0 streams 2 =="; TMM_PLAN_MARGIN=$mg TMM_PLAN_KC0=$kc0 $D hostone N N 10000 10000 10000 0 1 2 4 2>&1 | grep -E "HOST|run 3"
done; done
echo "== margin 1.5 streams 4 =="; TMM_PLAN_MARGIN=1.5 $D hostone N N 10000 10000 10000 0 1 4 4 2>&1 | grep -E "HOST|run 3"
echo "== TRACE margin 1.5 streams 2 =="; TMM_TRACE=1 TMM_PLAN_MARGIN=1.5 $D hostone N N 10000 10000 10000 0 1 2 2 2>&1 | grep -E "trace|run" | tail -45
echo "== TRACE margin 1.3 streams 2 =="; TMM_TRACE=1 TMM_PLAN_MARGIN=1.3 $D hostone N N 10000 10000 10000 0 1 2 2 2>&1 | grep -E "trace|run" | tail -52

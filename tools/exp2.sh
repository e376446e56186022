#!/bin/bash
cd $GRAFT_REPO_ROOT
D=./build/devtest
for mg in 0.9 1.0 1.1 1.3 1.5; do for kc0 in 256 512; do for gr in 1.25 1.5 2.0; do
echo "== margin $mg kc0 $kc0 growth $gr =="; TMM_PLAN_GROWTH=$gr TMM_PLAN_MARGIN=$mg TMM_PLAN_KC0=$kc0 $D hostone N N 10000 10000 10000 0 1 2 4 2>&1 | grep -E "HOST|run 3"
done; done; done
echo "== TRACE margin 1.1 kc0 512 gr 1.5 =="; TMM_TRACE=1 TMM_PLAN_MARGIN=1.1 TMM_PLAN_KC0=512 TMM_PLAN_GROWTH=1.5 $D hostone N N 10000 10000 10000 0 1 2 2 2>&1 | grep -E "trace|run" | tail -40

#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
echo "== pytest gpu =="; timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -5
for sp in 1 2 3 4; do
echo "== P1SPLIT $sp =="; TMM_PLAN_P1SPLIT=$sp python tools/e2e.py --reps 6 2>&1 | grep -E "E2E"
done
echo "== P1SPLIT 2 margin 1.15 =="; TMM_PLAN_P1SPLIT=2 TMM_PLAN_MARGIN=1.15 python tools/e2e.py --reps 6 2>&1 | grep -E "E2E"
echo "== P1SPLIT 4 margin 1.15 =="; TMM_PLAN_P1SPLIT=4 TMM_PLAN_MARGIN=1.15 python tools/e2e.py --reps 6 2>&1 | grep -E "E2E"
echo "== P1SPLIT 4 kc0 384 =="; TMM_PLAN_P1SPLIT=4 TMM_PLAN_KC0=384 python tools/e2e.py --reps 6 2>&1 | grep -E "E2E"
echo "== trace default =="; python tools/e2e.py --reps 3 --trace 2>&1 | grep -E "trace|run|E2E" | tail -64
} 2>&1 | tee gpurun_out/exp_e2e2.txt

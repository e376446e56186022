#!/bin/bash
# One GPU-box pass: parity suite, smoke, both bench arms, ncu launch list and full captures of the DGEMM / ZGEMM kernels.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/box.txt; nproc >> gpurun_out/box.txt; free -g >> gpurun_out/box.txt
echo "== pytest gpu =="; timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
echo "== smoke =="; timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 | tee gpurun_out/smoke.txt
echo "== bench reference =="; timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_ref.json
echo "== bench ours =="; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.json
echo "== ncu launch list =="
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log
echo "== ncu full dgemm =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dgemm_kernel -s 4 -c 1 -o gpurun_out/prof_dgemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
echo "== ncu full zgemm =="
timeout 900 ncu --set full --clock-control none --import-source on -k regex:zgemm_kernel -s 2 -c 1 -o gpurun_out/prof_zgemm python tools/kbench.py z1 > gpurun_out/ncu_full_z.log 2>&1
tail -2 gpurun_out/ncu_full_z.log
ls -la gpurun_out

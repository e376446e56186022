#!/bin/bash
# Last GPU call of round 1 (one B200, ~4 min of box time): the whole -m gpu suite (incl. the new app / handle / re-pitch tests),
# the C++ miniapps (this repository's and the reference's own source) at the headline size, and one bench line.
mkdir -p gpurun_out
OUT=gpurun_out/r1_final.txt
{ nvidia-smi -L; date; } > $OUT
echo "== pytest -m gpu ==" >> $OUT
( timeout 200 python -m pytest tests -m gpu -q --timeout 150 2>&1 | tail -40 ) >> $OUT
echo "== smoke ==" >> $OUT
( timeout 60 python __graft_entry__.py smoke 2>&1 | tail -3 ) >> $OUT
echo "== bin/multiply 10000^3 ==" >> $OUT
( timeout 30 bin/multiply -m 10000 -n 10000 -k 10000 -r 3 2>&1 | tail -11 ) >> $OUT
echo "== bin/ref-multiply 10000^3 (reference source, this library) ==" >> $OUT
( timeout 30 bin/ref-multiply -m 10000 -n 10000 -k 10000 -r 3 2>&1 | tail -9 ) >> $OUT
echo "== bench ==" >> $OUT
( timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 ) >> $OUT
date >> $OUT
tail -60 $OUT

#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-z}.log
: > $OUT
echo "== pytest -m gpu (all)" >> $OUT
timeout -s KILL 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 >> $OUT
echo "== smoke" >> $OUT
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -5 >> $OUT
cat $OUT

#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_m.log
: > $OUT
echo "== pytest -m gpu (all)" >> $OUT
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q >> $OUT 2>&1
echo "rc=$?" >> $OUT
echo "== smoke" >> $OUT
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" >> $OUT 2>&1
echo "rc=$?" >> $OUT
echo "== sanitizer (chain: memcheck racecheck synccheck; others: synccheck)" >> $OUT
SAN_TIMEOUT=500 TOOLS="memcheck racecheck synccheck" bash tools/sanitize.sh >> $OUT 2>&1
tail -60 $OUT

#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/r02_probe_${1:-w}.log
: > $OUT
echo "== pytest -m gpu (f5 + fullsize)" >> $OUT
timeout -s KILL 1200 python -m pytest tests/test_gpu_f5.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -8 >> $OUT
echo "== attn_time" >> $OUT
timeout -s KILL 300 python tools/r02_probe.py attn_time 2>&1 >> $OUT
echo "== bench default" >> $OUT
timeout -s KILL 900 python bench.py > gpurun_out/bench_w.json 2> gpurun_out/bench_w.err
echo "rc=$?" >> $OUT
tail -c 6000 gpurun_out/bench_w.json >> $OUT
cat $OUT

#!/bin/bash
# CTA-pair GEMM vs one-CTA GEMM on the same box; short timeouts so that a deadlock cannot hold the box
OUT=gpurun_out/test_2sm_${1:-x}.log
: > $OUT
echo "== conv tests forced 2SM" >> $OUT
B200TTS_2SM=1 timeout -s KILL 120 python -m pytest tests/test_gpu_bigvgan.py -m gpu -x -q -k "conv1d_tcgen05 or conv_transpose" >> $OUT 2>&1
echo "rc=$?" >> $OUT
echo "== bench_gemm forced 2SM" >> $OUT
B200TTS_2SM=1 timeout -s KILL 120 python tools/bench_gemm.py vgan.s0 vgan.s1 vgan.s2 bat. res.s1 res.s2 >> $OUT 2>&1
echo "rc=$?" >> $OUT
echo "== bench_gemm 1SM" >> $OUT
B200TTS_2SM=0 timeout -s KILL 120 python tools/bench_gemm.py vgan. bat. res. dit. >> $OUT 2>&1
echo "== full tests auto(-1)" >> $OUT
B200TTS_2SM=-1 timeout -s KILL 400 python -m pytest tests -m gpu -x -q >> $OUT 2>&1
echo "rc=$?" >> $OUT

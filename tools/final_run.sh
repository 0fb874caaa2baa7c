#!/bin/bash
# End-of-round measurement pass on the GPU box: full GPU test suite, smoke, every bench workload.
set -x
TAG=${1:-r01final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 2>gpurun_out/bench_err_$TAG.log > gpurun_out/bench_bigvgan_$TAG.json
timeout 300 python bench.py --workload f5 --steps 5 --warmup 3 2>>gpurun_out/bench_err_$TAG.log > gpurun_out/bench_f5_$TAG.json
timeout 300 python bench.py --workload pipeline --steps 3 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_err_$TAG.log > gpurun_out/bench_pipeline_$TAG.json
timeout 300 python bench.py --workload indextts_gpt --steps 3 --warmup 3 2>>gpurun_out/bench_err_$TAG.log > gpurun_out/bench_igpt_$TAG.json
timeout 300 python bench.py --workload indextts_vocoder --steps 10 --warmup 3 --no-cpu-baseline 2>>gpurun_out/bench_err_$TAG.log > gpurun_out/bench_ivgan_$TAG.json
for f in bigvgan f5 pipeline igpt ivgan; do cut -c1-230 gpurun_out/bench_${f}_$TAG.json; done
tail -3 gpurun_out/bench_err_$TAG.log
# launch lists of the same commands (cold-cache, serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bigvgan_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_igpt_$TAG.csv \
    python tools/prof_f5.py --what igpt --steps 6 > /dev/null 2>&1
ls -la gpurun_out/launches_*_$TAG.csv

#!/bin/bash
# end-of-round pass on HEAD: the whole GPU suite, the default bench line, smoke
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r03_pytest_gpu.log 2>&1; tail -3 gpurun_out/r03_pytest_gpu.log
timeout 200 python bench.py --steps 5 --warmup 3 > gpurun_out/r03_bench_config4.json 2> gpurun_out/r03_bench_config4.err; cut -c1-400 gpurun_out/r03_bench_config4.json; tail -2 gpurun_out/r03_bench_config4.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

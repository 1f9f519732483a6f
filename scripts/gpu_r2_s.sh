#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
( time timeout 1500 python bench.py --workload conformer_large_lm_beam60_16x30s --steps 3 --warmup 3 > gpurun_out/r2s_bench_lm60.json 2> gpurun_out/r2s_bench_lm60.err ) 2>&1 | grep real
echo "exit $?"; cut -c1-2500 gpurun_out/r2s_bench_lm60.json; tail -4 gpurun_out/r2s_bench_lm60.err | cut -c1-400

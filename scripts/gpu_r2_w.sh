#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1 | grep -E "passed|failed|error|real"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench.json 2>/dev/null; cut -c1-250 gpurun_out/r2w_bench.json
ESPB_DWCONV_V1=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench_v1.json 2>/dev/null; cut -c1-250 gpurun_out/r2w_bench_v1.json
ESPB_PDL=0 timeout 600 python bench.py --steps 1 --warmup 2 --no-cpu-baseline --trace > /dev/null 2> gpurun_out/r2w_trace_pdl_off.txt; grep "\[trace\]" gpurun_out/r2w_trace_pdl_off.txt | head -24

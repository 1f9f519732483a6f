#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1 | grep -E "passed|failed|error|real"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -1
ESPB_PDL=0 timeout 600 python bench.py --steps 1 --warmup 2 --no-cpu-baseline --trace > /dev/null 2> gpurun_out/r2x_trace_pdl_off.txt; grep "\[trace\]" gpurun_out/r2x_trace_pdl_off.txt | grep -E "device activities|ctc_score|ctc_advance|rows_topk"
( time timeout 900 python bench.py > gpurun_out/r2x_bench_default.json 2> gpurun_out/r2x_bench_default.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/r2x_bench_default.json

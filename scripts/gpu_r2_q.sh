#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
ESPB_TRACE_DUMP=30 ESPB_PDL=0 timeout 600 python bench.py --steps 1 --warmup 2 --no-cpu-baseline --trace > gpurun_out/r2q_trace.json 2> gpurun_out/r2q_trace_pdl_off.txt; grep "\[timeline\]" gpurun_out/r2q_trace_pdl_off.txt
ESPB_TRACE_DUMP=30 timeout 600 python bench.py --steps 1 --warmup 2 --no-cpu-baseline --trace > gpurun_out/r2q_trace2.json 2> gpurun_out/r2q_trace_pdl_on.txt; grep "\[timeline\]" gpurun_out/r2q_trace_pdl_on.txt

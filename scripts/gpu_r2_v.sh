#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
python scripts/kernel_microbench.py logsoftmax 9 2>&1 | grep kernel_microbench
ESPB_LOGSOFTMAX_3PASS=1 python scripts/kernel_microbench.py logsoftmax 9 2>&1 | grep kernel_microbench
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 ) 2>&1 | grep -E "passed|failed|error|real"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench.json 2>/dev/null; cut -c1-250 gpurun_out/r2v_bench.json
ESPB_LOGSOFTMAX_3PASS=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench_3pass.json 2>/dev/null; cut -c1-250 gpurun_out/r2v_bench_3pass.json

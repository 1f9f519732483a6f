#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu -x -p no:cacheprovider --timeout 120 2>&1 | tail -3
ATTN_BENCH_ONLY=fused timeout 300 python scripts/attn_microbench.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_scorer_interface.py -q -m gpu -x -p no:cacheprovider --timeout 300 > gpurun_out/r2c_scorer.log 2>&1; echo "scorer iface exit $?"; tail -25 gpurun_out/r2c_scorer.log | cut -c1-300

#!/bin/bash
# What the driver runs at round end: build, the whole -m gpu suite in one process, smoke().
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 2400 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/full_gpu_tests.log 2>&1
echo "== full gpu suite -> exit $?"; tail -6 gpurun_out/full_gpu_tests.log | cut -c1-250
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

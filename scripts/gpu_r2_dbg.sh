#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
which compute-sanitizer
timeout 300 compute-sanitizer --tool memcheck --print-limit 8 --launch-timeout 60 python -m pytest "tests/test_gpu_attention.py" -q -m gpu -x -p no:cacheprovider -k "1-1-37-lens0 and False" > gpurun_out/r2_dbg_plain.log 2>&1
echo "plain exit $?"; grep -vE "^\s*$" gpurun_out/r2_dbg_plain.log | grep -E "=========|passed|failed" | head -40
timeout 300 compute-sanitizer --tool memcheck --print-limit 8 --launch-timeout 60 python -m pytest "tests/test_gpu_attention.py" -q -m gpu -x -p no:cacheprovider -k "1-1-37-lens0 and True" > gpurun_out/r2_dbg_rel.log 2>&1
echo "relpos exit $?"; grep -vE "^\s*$" gpurun_out/r2_dbg_rel.log | grep -E "=========|passed|failed" | head -40

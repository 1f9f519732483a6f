#!/bin/bash
# Runs the -m gpu tests in separate processes (a CUDA fault in one must not hide the others); logs to gpurun_out/.
# usage: scripts/gpu_tests.sh [stage ...]   stages: gemm_simt gemm_tc gemm_tc2 pipe_simt pipe_tc pipe_tc2   (default: all)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
stages=${@:-gemm_simt gemm_tc gemm_tc2 pipe_simt pipe_tc pipe_tc2}
rc=0
for st in $stages; do
  case $st in
    gemm_simt) mode=simt; file=tests/test_gpu_gemm.py;;
    gemm_tc)   mode=tc;   file=tests/test_gpu_gemm.py;;
    gemm_tc2)  mode=tc2;  file=tests/test_gpu_gemm.py;;
    pipe_tc2)  mode=tc2;  file=tests/test_gpu_pipeline.py;;
    pipe_simt) mode=simt; file=tests/test_gpu_pipeline.py;;
    pipe_tc)   mode=tc;   file=tests/test_gpu_pipeline.py;;
    *) echo "unknown stage $st"; continue;;
  esac
  ESPB_TEST_GEMM_MODES=$mode ESPNET_B200_GEMM=$mode timeout 1200 python -m pytest $file -q -m gpu -s -p no:cacheprovider --timeout 300 > gpurun_out/$st.log 2>&1
  r=$?
  echo "== $st -> exit $r"; grep -E "passed|failed|error" gpurun_out/$st.log | tail -3; grep -E "^(FAILED|ERROR)|max abs err|Error|error:" gpurun_out/$st.log | cut -c1-300 | head -40
  [ $r -ne 0 ] && rc=$r
done
exit $rc

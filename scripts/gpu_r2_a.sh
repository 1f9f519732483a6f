#!/bin/bash
# Round 2, first GPU contact: default-GEMM parity suites (incl. the benchmarked-configuration and TransformerEncoder rows) + the default bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for f in test_gpu_gemm test_gpu_pipeline test_gpu_large test_gpu_zz_next; do
  ESPB_TEST_GEMM_MODES=tc2 timeout 1500 python -m pytest tests/$f.py -q -m gpu -s -p no:cacheprovider --timeout 600 > gpurun_out/r2a_$f.log 2>&1
  echo "== $f -> exit $?"; grep -E "passed|failed|error" gpurun_out/r2a_$f.log | tail -2; grep -E "^(FAILED|ERROR)|max abs err|Error|error:|steps=" gpurun_out/r2a_$f.log | cut -c1-260 | head -40
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench exit $?"; cut -c1-3000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err | cut -c1-600
nproc

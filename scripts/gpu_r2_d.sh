#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for f in test_gpu_gemm test_gpu_attention test_gpu_pipeline test_scorer_interface; do
  ESPB_TEST_GEMM_MODES=tc2,tc timeout 1500 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/r2d_$f.log 2>&1
  echo "== $f -> exit $?"; grep -E "passed|failed|error" gpurun_out/r2d_$f.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2d_$f.log | cut -c1-200 | head
done
ATTN_BENCH_ONLY=fused timeout 300 python scripts/attn_microbench.py 2>&1 | tail -4
timeout 300 python scripts/gemm_microbench.py 2>&1 | tail -12
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench exit $?"; cut -c1-600 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err | cut -c1-300

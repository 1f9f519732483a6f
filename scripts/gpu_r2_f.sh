#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
for f in test_gpu_pipeline test_gpu_large; do
  ESPB_TEST_GEMM_MODES=tc2 timeout 1500 python -m pytest tests/$f.py -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/r2f_$f.log 2>&1
  echo "== $f -> exit $?"; grep -E "passed|failed|error" gpurun_out/r2f_$f.log | tail -2; grep -E "^(FAILED|ERROR)" gpurun_out/r2f_$f.log | cut -c1-200 | head
done
timeout 300 python scripts/kernel_microbench.py srcattn 5 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench exit $?"; cut -c1-400 gpurun_out/r2f_bench.json

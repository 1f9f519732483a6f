#!/bin/bash
# fused attention bring-up: unit test vs fp64, encoder-level parity, microbench, bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 600 python -m pytest tests/test_gpu_attention.py -q -m gpu -s -x -p no:cacheprovider --timeout 120 > gpurun_out/r2b_attn.log 2>&1
echo "== attention unit -> exit $?"; grep -E "passed|failed|error" gpurun_out/r2b_attn.log | tail -2; grep -E "max abs err|Error|error|rror" gpurun_out/r2b_attn.log | cut -c1-200 | head -40
if grep -q "failed\|rror" gpurun_out/r2b_attn.log; then tail -40 gpurun_out/r2b_attn.log; exit 1; fi
timeout 300 python scripts/attn_microbench.py > gpurun_out/r2b_attn_bench.txt 2>&1; cat gpurun_out/r2b_attn_bench.txt | tail -8
for f in test_gpu_pipeline test_gpu_large test_gpu_zz_next; do
  ESPB_TEST_GEMM_MODES=tc2 timeout 1500 python -m pytest tests/$f.py -q -m gpu -s -p no:cacheprovider --timeout 600 > gpurun_out/r2b_$f.log 2>&1
  echo "== $f -> exit $?"; grep -E "passed|failed|error" gpurun_out/r2b_$f.log | tail -2; grep -E "^(FAILED|ERROR)|encoder max abs err|Error|error:" gpurun_out/r2b_$f.log | cut -c1-260 | head -20
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench exit $?"; cut -c1-700 gpurun_out/r2b_bench.json; tail -3 gpurun_out/r2b_bench.err | cut -c1-300

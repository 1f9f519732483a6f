#!/bin/bash
# band GEMM: 32x32 blocks outside the rel-pos band are not stored; refactored 1-CTA epilogue
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_attention.py tests/test_gpu_large.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "encoder or long_utterance or config1" 2>&1 | tail -3
ATTN_BENCH_ONLY=fused timeout 300 python scripts/attn_microbench.py 64 937 5 2>&1 | grep -i -E "band|fused|bd" | head -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; cut -c1-260 gpurun_out/r2t_bench.json

#!/bin/bash
# cross-attention with 4 warps / 32-frame tiles per block (4 blocks per SM) vs 8 warps / 64-frame tiles (2 per SM)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
ESPB_SRC_ATTN_NW=4 timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_large.py -x -q -m gpu -k "beam or config1 or long_utterance or large or golden" 2>&1 | tail -3
python scripts/kernel_microbench.py srcattn 9 2>&1 | grep kernel_microbench
ESPB_SRC_ATTN_NW=4 python scripts/kernel_microbench.py srcattn 9 2>&1 | grep kernel_microbench
ESPB_SRC_ATTN_NW=4 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_nw4.json 2>/dev/null; cut -c1-250 gpurun_out/r2u_bench_nw4.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_nw8.json 2>/dev/null; cut -c1-250 gpurun_out/r2u_bench_nw8.json
ESPB_SRC_ATTN_NW=4 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_nw4b.json 2>/dev/null; cut -c1-250 gpurun_out/r2u_bench_nw4b.json

#!/bin/bash
# streaming beam search on the GPU; split-only EW16 dispatch: microbench + bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_streaming_search.py tests/test_streaming_encoder.py tests/test_scorer_interface.py tests/test_gpu_gemm.py tests/test_capi_symbols.py -x -q -m gpu 2>&1 | tail -6
python scripts/gemm_enc_microbench.py 2>&1 | grep gemm_enc
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; cut -c1-260 gpurun_out/r2p_bench.json

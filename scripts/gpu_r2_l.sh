#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_streaming_encoder.py tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider --timeout 300 -k "streaming or cli" > gpurun_out/r2l_stream.log 2>&1; echo "streaming tests exit $?"; tail -15 gpurun_out/r2l_stream.log | cut -c1-250
timeout 600 python bench.py --workload streaming_cbconformer_128x40ms --steps 3 --warmup 2 > gpurun_out/r2l_bench_stream.json 2> gpurun_out/r2l_bench_stream.err; echo "stream bench exit $?"; cut -c1-900 gpurun_out/r2l_bench_stream.json; tail -3 gpurun_out/r2l_bench_stream.err | cut -c1-300

#!/bin/bash
# BASELINE configs[3]: 1024 live streams over 8 GPUs (128 per GPU), 40-ms pushes
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --workload streaming_cbconformer_128x40ms > gpurun_out/r2_bench_streaming_n8.json 2> gpurun_out/r2_bench_streaming_n8.err
echo "exit $?"; grep '^{' gpurun_out/r2_bench_streaming_n8.json | cut -c1-900; tail -2 gpurun_out/r2_bench_streaming_n8.err | cut -c1-200

#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
echo "exit $?"; grep '^{' gpurun_out/r2_bench_n2_final.json | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['config'].get('extra'), d['clocks'])"
tail -2 gpurun_out/r2_bench_n2_final.err | cut -c1-200

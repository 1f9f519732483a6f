#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "bench n8 exit $?"; python - <<'PY'
import json
for ln in open('gpurun_out/r2_bench_n8.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['ms_per_step'], d['config'].get('extra'), d['clocks'])
PY
tail -2 gpurun_out/r2_bench_n8.err | cut -c1-200

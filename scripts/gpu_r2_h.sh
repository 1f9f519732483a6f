#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python -m pytest tests/test_lm_fusion.py -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/r2h_lm.log 2>&1; echo "lm exit $?"; tail -5 gpurun_out/r2h_lm.log | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err
echo "bench n2 exit $?"; python - <<'PY'
import json
for ln in open('gpurun_out/r2h_bench_n2.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'], d['config'].get('extra'))
PY
tail -3 gpurun_out/r2h_bench_n2.err | cut -c1-300

#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python - <<'PY'
import json
for ln in open('gpurun_out/r2i_bench.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'])
PY

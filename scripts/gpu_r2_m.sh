#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
ESPB_TEST_GEMM_MODES=tc2 timeout 900 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider --timeout 300 -k "beam" > gpurun_out/r2m_beam.log 2>&1; echo "beam tests exit $?"; tail -4 gpurun_out/r2m_beam.log | cut -c1-250
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench exit $?"
python - <<'PY'
import json
for ln in open('gpurun_out/r2m_bench.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d.get('parity_check'), d.get('cpu_baseline',{}).get('value'))
        for o in d['roofline']['other_kernels']: print(o)
        print(d['roofline']['achieved'], d['roofline']['frac'])
PY
bash scripts/gpu_launchlist.sh conformer_large_joint_64x30s r02 2>&1 | tail -32

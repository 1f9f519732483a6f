#!/bin/bash
# Next scope row: first GPU contact of the TransformerEncoder path.
#   1. opt-in parity tests (reference fixtures + oracle), 2. the config-5-shaped bench workload (attention-only, beam 5).
# usage (on the GPU box): bash scripts/gpu_next.sh
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
ESPB_TEST_NEXT=1 timeout 900 python -m pytest tests/test_gpu_zz_next.py -q -m gpu -s -p no:cacheprovider > gpurun_out/next_tests.log 2>&1
echo "next tests exit $?"; tail -15 gpurun_out/next_tests.log
timeout 900 python bench.py --workload transformer_24l1024_att_64x30s --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_next.json 2> gpurun_out/bench_next.err
echo "bench exit $?"; cut -c1-400 gpurun_out/bench_next.json; tail -3 gpurun_out/bench_next.err | cut -c1-400

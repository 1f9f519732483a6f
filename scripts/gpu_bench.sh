#!/bin/bash
# bench runs with logs into gpurun_out/ (usage: scripts/gpu_bench.sh <workload> [steps] [warmup])
mkdir -p gpurun_out
wl=${1:-conformer_large_joint_64x30s}; steps=${2:-3}; warm=${3:-3}; mode=${4:-tc}; extra=${5:-}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
ESPNET_B200_GEMM=$mode timeout 900 python bench.py --workload $wl --steps $steps --warmup $warm $extra > gpurun_out/bench_${wl}_$mode.json 2> gpurun_out/bench_${wl}_$mode.err
echo "bench $wl $mode exit $?"; tail -c 3000 gpurun_out/bench_${wl}_$mode.json; tail -3 gpurun_out/bench_${wl}_$mode.err | cut -c1-3000

#!/bin/bash
# bench runs with logs into gpurun_out/ (usage: scripts/gpu_bench.sh <workload> [steps] [warmup])
mkdir -p gpurun_out
wl=${1:-conformer_large_joint_64x30s}; steps=${2:-3}; warm=${3:-3}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python bench.py --workload $wl --steps $steps --warmup $warm > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err
echo "bench $wl exit $?"; tail -c 3000 gpurun_out/bench_$wl.json; tail -5 gpurun_out/bench_$wl.err

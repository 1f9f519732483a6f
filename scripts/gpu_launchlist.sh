#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of one bench step. usage: scripts/gpu_launchlist.sh <workload> <tag>
mkdir -p gpurun_out
wl=${1:-conformer_large_joint_64x30s}; tag=${2:-r01b}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --profile-one-step > gpurun_out/prof_bench_${tag}.log 2>&1
echo "launch list exit $?"
python scripts/summarize_launches.py gpurun_out/launches_${tag}.csv > gpurun_out/launch_summary_${tag}.txt 2>&1; head -40 gpurun_out/launch_summary_${tag}.txt

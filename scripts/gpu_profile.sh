#!/bin/bash
# ncu evidence (B200_PROFILING.md recipe): launch list of one bench step + one --set full capture of the dominant kernel.
# usage: scripts/gpu_profile.sh <workload> <tag>
mkdir -p gpurun_out
wl=${1:-conformer_large_joint_32x15s}; tag=${2:-r01}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
# every launch with its device time (cold cache, serialised): compare SHARES, not absolutes
timeout 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --profile-one-step > gpurun_out/prof_bench_${tag}.log 2>&1
echo "launch list exit $?"
python scripts/summarize_launches.py gpurun_out/launches_${tag}.csv > gpurun_out/launch_summary_${tag}.txt 2>&1; head -40 gpurun_out/launch_summary_${tag}.txt
# the top kernel once, full set (about 40 replays per launch): 3 launches of the BN=256 GEMM after warm-up
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 40 -c 3 -o gpurun_out/gemm_${tag} -f \
    python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --profile-one-step > gpurun_out/prof_full_${tag}.log 2>&1
echo "full capture exit $?"
ncu -i gpurun_out/gemm_${tag}.ncu-rep --page raw --csv > gpurun_out/gemm_${tag}_raw.csv 2>/dev/null
grep -E 'Kernel Name|dram__bytes_(read|write)\.sum,|gpu__time_duration\.sum|sm__pipe_tensor|sm__inst_executed_pipe_tensor|launch__registers_per_thread|sm__warps_active|gpu__dram_throughput|lts__t_bytes\.sum,|sm__throughput' gpurun_out/gemm_${tag}_raw.csv | head -5 | cut -c1-300

#!/bin/bash
# ncu --set full capture of kernels matching a regex inside one bench step. usage: scripts/gpu_ncu_kernel.sh <regex> <tag> [skip] [count] [workload]
mkdir -p gpurun_out
re=$1; tag=$2; skip=${3:-0}; cnt=${4:-2}; wl=${5:-conformer_large_joint_64x30s}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 800 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$re -s $skip -c $cnt -o gpurun_out/$tag -f \
    python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --profile-one-step > gpurun_out/ncu_$tag.log 2>&1
echo "ncu $tag exit $?"
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_raw.csv

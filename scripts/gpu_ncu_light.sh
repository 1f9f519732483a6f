#!/bin/bash
# light ncu capture (SpeedOfLight + memory sections) usage: scripts/gpu_ncu_light.sh <regex> <tag> [skip] [count]
mkdir -p gpurun_out
re=$1; tag=$2; skip=${3:-0}; cnt=${4:-4}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 ncu --profile-from-start off --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --clock-control none -k regex:$re -s $skip -c $cnt --csv --page raw --log-file gpurun_out/${tag}_raw.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --profile-one-step > gpurun_out/ncu_$tag.log 2>&1
echo "ncu $tag exit $?"
grep -v "^==" gpurun_out/${tag}_raw.csv > gpurun_out/${tag}_raw2.csv
python scripts/ncu_summary.py gpurun_out/${tag}_raw2.csv

#!/bin/bash
# One --set full ncu capture of ONE launch of one kernel from scripts/kernel_microbench.py.
# usage: scripts/gpu_ncu_micro.sh <softmax|srcattn> <kernel-regex> <tag>
mkdir -p gpurun_out
what=$1; rx=$2; tag=$3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 1 -c 1 -o gpurun_out/$tag -f python scripts/kernel_microbench.py $what 1 > gpurun_out/ncu_$tag.log 2>&1
echo "ncu $tag exit $?"
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_raw.csv > gpurun_out/${tag}_summary.txt 2>&1; head -60 gpurun_out/${tag}_summary.txt

#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
ATTN_BENCH_ONLY=fused timeout 900 ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 2 -c 2 -o gpurun_out/r02_ncu_flash_attn_v2 -f python scripts/attn_microbench.py 64 937 1 > gpurun_out/ncu_v2.log 2>&1
ncu -i gpurun_out/r02_ncu_flash_attn_v2.ncu-rep --page raw --csv > gpurun_out/r02_ncu_flash_attn_v2_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_ncu_flash_attn_v2.ncu-rep --page source --csv > gpurun_out/r02_ncu_flash_attn_v2_source.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02_ncu_flash_attn_v2_raw.csv > gpurun_out/r02_ncu_flash_attn_v2_summary.txt 2>&1; grep -E "==|time_duration|dram__bytes|tensor|issue_active|stalled" gpurun_out/r02_ncu_flash_attn_v2_summary.txt | cut -c1-150
ATTN_BENCH_ONLY=fused python scripts/attn_microbench.py 16 3000 3 2>&1 | tail -3
ESPB_PDL=0 timeout 600 python bench.py --trace --steps 1 --warmup 3 2> gpurun_out/r02_trace.txt > /dev/null; head -45 gpurun_out/r02_trace.txt | cut -c1-200

#!/bin/bash
# ncu --set full of ONE launch each of the fused attention kernel and of the bd band GEMM (scripts/attn_microbench.py, fused legs only)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for spec in "flash_attn_kernel r02_ncu_flash_attn" "gemm_tf32x3_2cta r02_ncu_bd_band_gemm"; do
  set -- $spec
  ATTN_BENCH_ONLY=fused timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s 1 -c 1 -o gpurun_out/$2 -f python scripts/attn_microbench.py 64 937 1 > gpurun_out/ncu_$2.log 2>&1
  echo "ncu $2 exit $?"
  ncu -i gpurun_out/$2.ncu-rep --page raw --csv > gpurun_out/$2_raw.csv 2>/dev/null
  ncu -i gpurun_out/$2.ncu-rep --page source --csv > gpurun_out/$2_source.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/$2_raw.csv > gpurun_out/$2_summary.txt 2>&1; head -45 gpurun_out/$2_summary.txt
done

#!/bin/bash
# End-of-round check on one box: whole -m gpu suite + smoke, ncu of the FFN w_1 GEMM (16-warp epilogue), default bench line, A/B of the band-block skip
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
( time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) 2>&1 | grep -E "passed|failed|error|real"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke OK')" 2>&1 | tail -2
tag=r02_ncu_gemm_2cta_ffn_w1_ew16
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3_2cta -s 1 -c 1 -o gpurun_out/$tag -f python scripts/gemm_enc_microbench.py > gpurun_out/ncu_$tag.log 2>&1
echo "ncu exit $?"
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_raw.csv > gpurun_out/${tag}_summary.txt 2>&1; head -12 gpurun_out/${tag}_summary.txt
cp gpurun_out/${tag}_summary.txt profiles/${tag}_summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_blocks.json 2>/dev/null; cut -c1-250 gpurun_out/final_bench_blocks.json
ESPB_GEMM_BAND_TILES_ONLY=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_tiles.json 2>/dev/null; cut -c1-250 gpurun_out/final_bench_tiles.json
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_blocks2.json 2>/dev/null; cut -c1-250 gpurun_out/final_bench_blocks2.json
( time timeout 900 python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err ) 2>&1 | grep real; cut -c1-300 gpurun_out/final_bench_default.json

#!/bin/bash
# 16-warp epilogue of the 256-column CTA-pair GEMM: correctness, A/B microbench + bench; configs[4] Transformer bench line; ncu of the small kernels
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_large.py tests/test_gpu_attention.py -x -q -m gpu 2>&1 | tail -4
python scripts/gemm_enc_microbench.py 2>&1 | grep gemm_enc
ESPB_GEMM_EW8=1 python scripts/gemm_enc_microbench.py 2>&1 | grep gemm_enc
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_ew16.json 2> gpurun_out/r2o_bench_ew16.err; cut -c1-260 gpurun_out/r2o_bench_ew16.json
ESPB_GEMM_EW8=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2o_bench_ew8.json 2> gpurun_out/r2o_bench_ew8.err; cut -c1-260 gpurun_out/r2o_bench_ew8.json
( time timeout 900 python bench.py --workload transformer_24l1024_att_64x30s --steps 3 --warmup 3 > gpurun_out/r2o_bench_tfm.json 2> gpurun_out/r2o_bench_tfm.err ) 2>&1 | grep real; cut -c1-1500 gpurun_out/r2o_bench_tfm.json; tail -3 gpurun_out/r2o_bench_tfm.err | cut -c1-300
bash scripts/gpu_ncu_light.sh 'layernorm_vec|glu_dwconv|conv1_relu|qu_qv|v_transpose|transpose_tv|utt_mvn|ctc_init_state' r02_ncu_small_encoder 0 12 2>&1 | tail -60 > gpurun_out/r2o_ncu_enc.txt; tail -5 gpurun_out/r2o_ncu_enc.txt
bash scripts/gpu_ncu_light.sh 'ctc_advance|ctc_score_cands|rows_topk|log_softmax_rows|beam_select|dec_self_attn64|dec_src_attn_flash|dec_embed|anc_update' r02_ncu_small_search 40 22 2>&1 | tail -80 > gpurun_out/r2o_ncu_search.txt; tail -5 gpurun_out/r2o_ncu_search.txt

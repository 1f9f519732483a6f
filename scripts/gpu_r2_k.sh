#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
ESPB_TEST_GEMM_MODES=tc2 timeout 1500 python -m pytest tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/r2k_pipe.log 2>&1
echo "== pipeline -> exit $?"; grep -E "passed|failed|error" gpurun_out/r2k_pipe.log | tail -2; grep -E "^(FAILED|ERROR)|log-mel max" gpurun_out/r2k_pipe.log | cut -c1-200 | head
timeout 300 python scripts/kernel_microbench.py frontend 5 2>&1 | tail -2
ESPB_STFT_V1=1 timeout 300 python scripts/kernel_microbench.py frontend 5 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_logmel_v2 -s 1 -c 1 -o gpurun_out/r02_ncu_stft_logmel_v2 -f python scripts/kernel_microbench.py frontend 1 > gpurun_out/ncu_stft.log 2>&1
ncu -i gpurun_out/r02_ncu_stft_logmel_v2.ncu-rep --page raw --csv > gpurun_out/r02_ncu_stft_logmel_v2_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r02_ncu_stft_logmel_v2_raw.csv > gpurun_out/r02_ncu_stft_logmel_v2_summary.txt 2>&1; head -40 gpurun_out/r02_ncu_stft_logmel_v2_summary.txt | cut -c1-150

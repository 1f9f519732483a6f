#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "long_utterance or beam_wider" -s 2>&1 | grep -E "utt |passed|failed|Error|assert" | head -20
timeout 400 python scripts/streaming_search_microbench.py 20 500 10 2>&1 | grep -E "streaming_search|Error|error" | head -5

#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_ref.json 2> gpurun_out/r2n_ref.err ) 2>&1 | grep real; cut -c1-700 gpurun_out/r2n_ref.json; tail -2 gpurun_out/r2n_ref.err | cut -c1-200

#!/bin/bash
# Runs the -m gpu tests file by file (a CUDA fault in one file must not hide the others); logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; }
rc=0
for spec in "$@"; do
  name=$(echo "$spec" | tr '/:[] ' '_____')
  timeout 900 python -m pytest $spec -q -m gpu -s -p no:cacheprovider --timeout 300 > gpurun_out/$name.log 2>&1
  r=$?
  echo "== $spec -> exit $r"; tail -n 25 gpurun_out/$name.log | cut -c1-400
  [ $r -ne 0 ] && rc=$r
done
exit $rc

#!/bin/bash
# runs each GPU test file in its own process (a CUDA fault is sticky per process) and keeps full logs
mkdir -p gpurun_out
for f in tests/test_gpu_*.py; do
  b=$(basename $f .py)
  timeout 1200 python -m pytest $f -m gpu -q --timeout 900 --maxfail=8 > gpurun_out/$b.log 2>&1
  echo "$b rc=$? $(tail -1 gpurun_out/$b.log)"
done

#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_gpu_sumcheck.py tests/test_gpu_kzg.py; do
  b=$(basename $f .py)
  timeout 700 python -m pytest $f -m gpu -q -s --timeout 600 --maxfail=10 > gpurun_out/$b.log 2>&1
  echo "$b rc=$? $(tail -1 gpurun_out/$b.log)"
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader -i 0
timeout 600 python tools/n4_bench.py > gpurun_out/r2_n4_bench_n1.jsonl 2> gpurun_out/n4_bench.err
echo "n4_bench rc=$?"; cut -c1-210 gpurun_out/r2_n4_bench_n1.jsonl; tail -3 gpurun_out/n4_bench.err

#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_gpu_sumcheck.py tests/test_gpu_kzg.py; do
  b=$(basename $f .py)
  timeout 700 python -m pytest $f -m gpu -q -s --timeout 600 --maxfail=10 > gpurun_out/$b.log 2>&1
  echo "$b rc=$? $(tail -1 gpurun_out/$b.log)"
done
timeout 600 python tools/n4_bench.py > gpurun_out/r2_n4_bench_n1.jsonl 2> gpurun_out/n4_bench.err
echo "n4_bench rc=$?"; cut -c1-200 gpurun_out/r2_n4_bench_n1.jsonl; tail -3 gpurun_out/n4_bench.err
bash tools/profile_n34.sh > gpurun_out/profile_n34.log 2>&1
echo "profile rc=$?"; ls -la gpurun_out/r2_ncu_full_h2c_raw.csv gpurun_out/r2_ncu_full_sumcheck_raw.csv

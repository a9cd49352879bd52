#!/bin/bash
# GPU run for the KZG rows + N4 measurements (no Python inside the timed prover calls) + launch list of the N4 tool
mkdir -p gpurun_out
for f in tests/test_gpu_kzg.py tests/test_gpu_ck_generate.py; do
  b=$(basename $f .py)
  timeout 700 python -m pytest $f -m gpu -q -s --timeout 600 --maxfail=10 > gpurun_out/$b.log 2>&1
  echo "$b rc=$? $(tail -1 gpurun_out/$b.log)"
done
timeout 600 python tools/n4_bench.py > gpurun_out/r2_n4_bench_n1.jsonl 2> gpurun_out/n4_bench.err
echo "n4_bench rc=$?"; cat gpurun_out/r2_n4_bench_n1.jsonl; tail -3 gpurun_out/n4_bench.err
timeout 300 python tools/config_benches.py --only ckgen > gpurun_out/r2_config_ckgen_n1.jsonl 2> gpurun_out/ckgen.err
cat gpurun_out/r2_config_ckgen_n1.jsonl | cut -c1-330
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_n4.csv \
    python tools/n4_bench.py --logn 20 > gpurun_out/n4_under_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2_launches_n4.csv

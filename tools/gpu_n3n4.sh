#!/bin/bash
# GPU run for the N3 / N4 rows: parity tests, the key-generation measurements and one bench run on the hash-to-curve key
mkdir -p gpurun_out
for f in tests/test_gpu_ck_generate.py tests/test_gpu_sumcheck.py; do
  b=$(basename $f .py)
  timeout 600 python -m pytest $f -m gpu -q -s --timeout 500 --maxfail=10 > gpurun_out/$b.log 2>&1
  echo "$b rc=$? $(tail -1 gpurun_out/$b.log)"
done
timeout 300 python tools/config_benches.py --only ckgen > gpurun_out/r2_config_ckgen_n1.jsonl 2> gpurun_out/ckgen.err
echo "ckgen rc=$?"; cat gpurun_out/r2_config_ckgen_n1.jsonl
timeout 400 python bench.py --steps 20 --warmup 3 --key from_label --no-cpu-baseline > gpurun_out/r2_bench_n1_key_from_label.json 2> gpurun_out/bench_key.err
echo "bench rc=$?"; cat gpurun_out/r2_bench_n1_key_from_label.json | cut -c1-900

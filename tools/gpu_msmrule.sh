#!/bin/bash
mkdir -p gpurun_out
for f in tests/test_gpu_msm.py tests/test_gpu_kzg.py tests/test_gpu_sumcheck.py tests/test_gpu_zz_fold_size_commit.py tests/test_gpu_spartan_chain.py tests/test_gpu_fold_pipeline.py; do
  b=$(basename $f .py)
  timeout 700 python -m pytest $f -m gpu -q --timeout 600 --maxfail=10 > gpurun_out/$b.log 2>&1
  echo "$b rc=$? $(tail -1 gpurun_out/$b.log)"
done
timeout 900 python tools/compress_bench.py > gpurun_out/r2_compress_bench_n1.jsonl 2> gpurun_out/compress_bench.err
echo "compress rc=$?"; cat gpurun_out/r2_compress_bench_n1.jsonl; tail -3 gpurun_out/compress_bench.err
timeout 600 python tools/n4_bench.py --only kzg > gpurun_out/r2_n4_bench_kzg_hybrid.jsonl 2> gpurun_out/n4_kzg.err
cut -c1-230 gpurun_out/r2_n4_bench_kzg_hybrid.jsonl
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1_after_msmrule.json 2> gpurun_out/bench3.err
echo "bench rc=$?"; cut -c1-260 gpurun_out/r2_bench_n1_after_msmrule.json

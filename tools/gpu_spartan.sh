#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spartan_chain.py -m gpu -q -s --timeout 500 > gpurun_out/test_gpu_spartan_chain.log 2>&1
echo "spartan rc=$? $(tail -1 gpurun_out/test_gpu_spartan_chain.log)"; grep -E "^E  " gpurun_out/test_gpu_spartan_chain.log | head -8
timeout 600 python -m pytest tests/test_gpu_dag_fold.py tests/test_gpu_fold_pipeline.py -m gpu -q --timeout 500 > gpurun_out/test_gpu_dag_fold.log 2>&1
echo "dag_fold rc=$? $(tail -1 gpurun_out/test_gpu_dag_fold.log)"
timeout 900 python tools/compress_bench.py > gpurun_out/r2_compress_bench_n1.jsonl 2> gpurun_out/compress_bench.err
echo "compress rc=$?"; cat gpurun_out/r2_compress_bench_n1.jsonl; tail -4 gpurun_out/compress_bench.err

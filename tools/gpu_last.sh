#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/compress_bench.py > gpurun_out/r2_compress_bench_n1_b.jsonl 2> gpurun_out/compress_bench.err
echo "compress rc=$?"; cat gpurun_out/r2_compress_bench_n1_b.jsonl
timeout 120 python tools/n4_bench.py --only kzg > gpurun_out/r2_n4_bench_kzg_hybrid_b.jsonl 2> gpurun_out/n4_kzg.err
cut -c1-200 gpurun_out/r2_n4_bench_kzg_hybrid_b.jsonl

#!/bin/bash
# GPU run: N4 parity after the IPA rework (weighted Pippenger passes over the fixed key) + measurements + launch list
mkdir -p gpurun_out
for f in tests/test_gpu_sumcheck.py tests/test_gpu_kzg.py; do
  b=$(basename $f .py)
  timeout 700 python -m pytest $f -m gpu -q -s --timeout 600 --maxfail=10 > gpurun_out/$b.log 2>&1
  echo "$b rc=$? $(tail -1 gpurun_out/$b.log)"
done
timeout 600 python tools/n4_bench.py > gpurun_out/r2_n4_bench_n1.jsonl 2> gpurun_out/n4_bench.err
echo "n4_bench rc=$?"; cat gpurun_out/r2_n4_bench_n1.jsonl; tail -3 gpurun_out/n4_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_n4.csv \
    python tools/n4_bench.py --logn 20 > gpurun_out/n4_under_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2_launches_n4.csv
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

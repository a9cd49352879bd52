#!/bin/bash
# final validation of the round: every GPU test file, smoke, the default bench line, and the launch list of the bench command
mkdir -p gpurun_out
bash tools/gpu_tests.sh
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r2_bench_n1_head.json 2> gpurun_out/bench_head.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_bench_n1_head.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_head.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/r2_bench_reference_head.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_bench_head.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r2_launches_bench_head.csv

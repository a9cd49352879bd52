#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one bench.py run, (2) --set full captures of the two hot kernels.
# Numbers printed by processes running under ncu are never bench values.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:poseidon_kernel -s 2 -c 1 -o gpurun_out/prof_poseidon -f \
    python tools/quick_bench.py poseidon8 > gpurun_out/ncu_poseidon.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel -s 2 -c 1 -o gpurun_out/prof_msm -f \
    python tools/quick_bench.py msm21 > gpurun_out/ncu_msm.log 2>&1
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# Round-2 ncu evidence for profiles/ (run under gpurun, one GPU).  Numbers printed by processes running under ncu are never
# bench values.  (1) launch list of one bench.py run; (2) --set full captures of the dominant kernel inside the real step and
# of the single-CTA kernels of the chain; (3) the Poseidon witness kernel at 2^20.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:msm_accumulate_kernel -s 14 -c 1 -o gpurun_out/r2_prof_accumulate -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_acc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"fold_challenge_kernel|msm_horner_kernel|fold_axpy_kernel" -s 12 -c 3 -o gpurun_out/r2_prof_chain -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_chain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:poseidon_kernel -s 1 -c 1 -o gpurun_out/r2_prof_witness -f \
    python tools/config_benches.py --only witness --logn 20 > gpurun_out/r2_ncu_witness.log 2>&1
# raw pages are exported here (the box has the same ncu); the reports themselves are too large to travel back together
for r in accumulate chain witness; do
  ncu -i gpurun_out/r2_prof_$r.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_${r}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/r2_prof_accumulate.ncu-rep --page source --csv > gpurun_out/r2_ncu_accumulate_source.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/r2_prof_chain.ncu-rep gpurun_out/r2_prof_witness.ncu-rep
du -sh gpurun_out

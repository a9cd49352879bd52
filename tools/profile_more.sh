#!/bin/bash
# ncu --set full captures of the remaining kernels (HBM-bound helpers, NTT, slot-witness kernel, fixed-base accumulate)
set -x
mkdir -p gpurun_out
for spec in "axpy_kernel hbm 2" "cross_term_kernel hbm 2" "spmv_kernel hbm 2" "ntt_tile_kernel hbm 3" "ntt_stage2_kernel hbm 6" ; do
  set -- $spec
  ncu --set full --clock-control none -k regex:$1 -s $3 -c 1 -o gpurun_out/prof_$1 -f \
      python tools/config_benches.py --only $2 > gpurun_out/ncu_$1.log 2>&1
done
ncu --set full --clock-control none -k regex:poseidon_warp_kernel -s 3 -c 1 -o gpurun_out/prof_poseidon_warp -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_pw.log 2>&1
ncu --set full --clock-control none -k regex:msm_accumulate_kernel -s 8 -c 1 -o gpurun_out/prof_msm_fixed -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_mf.log 2>&1
ncu --set full --clock-control none -k regex:msm_bucket_reduce_kernel -s 8 -c 1 -o gpurun_out/prof_bucket_reduce -f \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_br.log 2>&1
for r in gpurun_out/prof_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}_raw.csv 2>/dev/null; rm -f $r; done
ls -la gpurun_out/

#!/bin/bash
# ncu --set full captures of the two new throughput kernels (N3 map kernel, N4 sum-check round kernel); raw pages exported on the box.
# Numbers printed by processes running under ncu are never bench values.
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:h2c_kernel -s 1 -c 1 -o gpurun_out/r2_prof_h2c -f \
    python tools/config_benches.py --only ckgen --logn 19 > gpurun_out/r2_ncu_h2c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sc_round_kernel -s 21 -c 2 -o gpurun_out/r2_prof_sumcheck -f \
    python tools/n4_bench.py --only sumcheck --logn 21 > gpurun_out/r2_ncu_sumcheck.log 2>&1
for r in h2c sumcheck; do
  ncu -i gpurun_out/r2_prof_$r.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_${r}_raw.csv 2>/dev/null
done
ls -la gpurun_out/*.ncu-rep
rm -f gpurun_out/r2_prof_h2c.ncu-rep gpurun_out/r2_prof_sumcheck.ncu-rep

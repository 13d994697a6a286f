#!/bin/bash
# Round-2 evidence, second pass (one B200): ncu --set full of the kernel families the first pass did not reach.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --set full -k regex:"k_sc_" -c 8 -o gpurun_out/r2_sumcheck -f python tools/time_sumcheck.py 20 once > /dev/null 2>&1
$NCU --set full -k regex:"k_rep3_batch|k_rep3_masks|k_fr_rand|k_vec_lincomb" -c 6 -o gpurun_out/r2_misc -f python tools/run_misc_once.py > /dev/null 2>&1
$NCU --set full -k regex:"k_msm_digits|k_msm_scan|k_msm_final_sum|k_msm_accum2|k_msm_slice" -s 9 -c 10 -o gpurun_out/r2_msm_sort -f python tools/run_msm_once.py 20 0 2 > /dev/null 2>&1
$NCU --set full -k regex:"k_plonk|k_scan|k_poly_eval|k_r3" -c 24 -o gpurun_out/r2_plonk_all -f python tools/run_plonk_once.py 16 > /dev/null 2>&1
for r in sumcheck misc msm_sort plonk_all; do
  if [ -f gpurun_out/r2_$r.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/r2_$r.ncu-rep gpurun_out/r2_ncu_full_$r.csv
    rm -f gpurun_out/r2_$r.ncu-rep
  else echo "missing capture $r"; fi
done

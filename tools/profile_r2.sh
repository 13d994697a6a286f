#!/bin/bash
# Round-2 evidence run (one B200, under gpurun): launch list of a bench step, ncu --set full captures of the kernels
# the verdict names, compute-sanitizer over the CI-size GPU tests.  Outputs land in gpurun_out/ and are summarised
# into profiles/ by tools/ncu_summary.py / tools/kernel_shares.py on the CPU box.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. launch list of the default bench (plain proof, 2 timed steps; rep3 block skipped: multi-threaded contexts)
$NCU --metrics gpu__time_duration.sum -c 2500 --csv --log-file gpurun_out/r2_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-rep3 --no-verify > gpurun_out/r2_launches_bench.out 2>&1
# 2. full captures
$NCU --set full --import-source on -k regex:k_msm_accum0 -s 1 -c 1 -o gpurun_out/r2_accum0_g1 -f python tools/run_msm_once.py 20 0 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_msm_accum0 -s 1 -c 1 -o gpurun_out/r2_accum0_g2 -f python tools/run_msm_once.py 18 1 3 > /dev/null 2>&1
$NCU --set full -k regex:"k_msm_scatter|k_msm_reduce_seg|k_msm_digits|k_msm_scan|k_msm_accum1" -s 8 -c 6 -o gpurun_out/r2_msm_tails -f python tools/run_msm_once.py 20 0 2 > /dev/null 2>&1
CS_NTT_V2=1 $NCU --set full --import-source on -k regex:k_ntt_pass -s 12 -c 6 -o gpurun_out/r2_ntt_v2 -f python tools/time_ntt.py 20:1 > /dev/null 2>&1
CS_NTT_V2=0 $NCU --set full -k regex:k_ntt_pass -s 8 -c 4 -o gpurun_out/r2_ntt_v1 -f python tools/time_ntt.py 20:1 > /dev/null 2>&1
$NCU --set full -k regex:"k_spmv|k_rep3_local_mul|k_plain_mul_sub|k_rep3_masks" -c 6 -o gpurun_out/r2_witness_map -f python tools/time_rep3_local.py 18 > /dev/null 2>&1
$NCU --set full -k regex:"k_plonk_quotient|k_r3_quot" -c 2 -o gpurun_out/r2_plonk_quotient -f python tools/run_plonk_once.py 16 > /dev/null 2>&1
# summaries here (ncu -i needs no GPU); the .ncu-rep files are too large to travel back
for r in accum0_g1 accum0_g2 msm_tails ntt_v2 ntt_v1 witness_map plonk_quotient; do
  if [ -f gpurun_out/r2_$r.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/r2_$r.ncu-rep gpurun_out/r2_ncu_full_$r.csv
    if [ "$r" = "ntt_v2" ]; then ncu -i gpurun_out/r2_$r.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/r2_ncu_source_$r.csv; fi
    rm -f gpurun_out/r2_$r.ncu-rep
  else echo "missing capture $r"; fi
done
# 3. compute-sanitizer on the small GPU parity tests (memcheck + racecheck)
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "field_ops or share_kernels or ntt_small or msm_g1_tiny or groth16_multiplier2 or rep3_batch_vm_ops" > gpurun_out/r2_sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ntt_small or share_kernels" > gpurun_out/r2_sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck.log
ls -la gpurun_out | tail -20

#!/bin/bash
# sweeps the NTT tuning hooks; one JSON line per variant into gpurun_out/ntt_sweep.jsonl
mkdir -p gpurun_out
: > gpurun_out/ntt_sweep.jsonl
for tws in 0 1; do
  for thr in 512 256 128; do
    for minb in 1; do
      if [ $minb = 3 ] && [ $thr != 512 ]; then continue; fi
      echo "tws=$tws thr=$thr minb=$minb" >> gpurun_out/ntt_sweep.jsonl
      CS_NTT_TWS=$tws CS_NTT_THREADS=$thr CS_NTT_MINB=$minb timeout 300 python tools/time_ntt.py >> gpurun_out/ntt_sweep.jsonl 2>&1
    done
  done
done

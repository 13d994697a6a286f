#!/bin/bash
# Round-2 evidence, third pass (two B200s, one process): the fused Rep3 product + NVLink peer-store kernel under ncu.
set -u
mkdir -p gpurun_out
python tools/run_mul_vec_peer_once.py 22 4 > gpurun_out/r2_mul_vec_peer_timing.log 2>&1
ncu --clock-control none --set full -k regex:k_rep3_mul_vec_reshare -c 2 -o gpurun_out/r2_mul_vec_peer -f python tools/run_mul_vec_peer_once.py 22 1 > /dev/null 2>&1
if [ -f gpurun_out/r2_mul_vec_peer.ncu-rep ]; then
  python tools/ncu_summary.py gpurun_out/r2_mul_vec_peer.ncu-rep gpurun_out/r2_ncu_full_mul_vec_peer.csv
  ncu -i gpurun_out/r2_mul_vec_peer.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
for li, vals in enumerate(rows[2:]):
    for h, u, v in zip(hdr, rows[1], vals):
        if 'nvl' in h.lower() or 'pcie' in h.lower() or h in ('lts__t_sectors_srcunit_tex_aperture_peer_op_write.sum', 'lts__t_sectors_aperture_peer.sum'):
            print(li, h, u, v)
" > gpurun_out/r2_mul_vec_peer_nvlink_metrics.txt
  rm -f gpurun_out/r2_mul_vec_peer.ncu-rep
else echo "missing capture" ; fi
ncu --clock-control none --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum -k regex:k_rep3_mul_vec_reshare -c 2 --csv --log-file gpurun_out/r2_mul_vec_peer_nvl.csv python tools/run_mul_vec_peer_once.py 22 1 > /dev/null 2>&1
cat gpurun_out/r2_mul_vec_peer_timing.log

#!/bin/bash
# ncu launch list of the unmodified reference FCM (config 3) for a per-stage comparison
python - <<'PY'
import sys; sys.path.insert(0,'.')
import bench_fcm as fcm_bench
p,f = fcm_bench.inputs(); p.tofile('/tmp/p.bin'); f.tofile('/tmp/f.bin')
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 80 --csv --log-file gpurun_out/ref_fcm_launches.csv oracle/_ref/ref_fcm time peskin3 500000 128 128 1.0 1e-3 1.0 0.01 3 4 1 /tmp/p.bin /tmp/f.bin > /dev/null 2>&1

#!/usr/bin/env bash
# GPU call AB (one B200): A/B of the k_voxelize shape that runs beside the map update (half: 512 x 2 at 64 registers; lean: 1024 x 1 at 48)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 0 1 0 1; do
  LIMU_VX_BESIDE=$v timeout 600 python bench.py --steps 150 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/ab_bench_$v.json 2> gpurun_out/ab_bench_$v.err
  grep -h '^{' gpurun_out/ab_bench_$v.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('beside shape $v', round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'))
"
done
LIMU_VX_BESIDE=1 LIMU_LIB=lidar-imu-slam_b200/build/liblimu_phase.so LIMU_SPECULATE=0 timeout 300 python tools/frame_phase_timing.py 2>&1 | grep -A16 "pipelined path" | tee gpurun_out/ab_phase_lean.txt

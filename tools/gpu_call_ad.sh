#!/usr/bin/env bash
# GPU call AD (one B200): programmatic dependent launch of k_voxelize behind the loop kernel (A/B with LIMU_NO_PDL=1), GPU suite, timeline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-ad}"
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
for v in pdl nopdl pdl nopdl; do
  if [ "$v" = "nopdl" ]; then export LIMU_NO_PDL=1; else unset LIMU_NO_PDL; fi
  timeout 600 python bench.py --steps 150 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/${T}_bench_$v.json 2> gpurun_out/${T}_bench_$v.err
  grep -h '^{' gpurun_out/${T}_bench_$v.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$v', round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'))
"
done
unset LIMU_NO_PDL
LIMU_LIB=lidar-imu-slam_b200/build/liblimu_phase.so LIMU_SPECULATE=0 timeout 300 python tools/frame_phase_timing.py 2>&1 | grep -A16 "pipelined path" | tee gpurun_out/${T}_phase.txt

#!/usr/bin/env bash
# GPU call Q/R (one B200): pipelined path; k_voxelize variants; GPU suite, bench, phase timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-q}"
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/${T}_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 2 ) > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err; echo "bench20 rc=$?"
( time timeout 900 python bench.py --steps 150 --warmup 5 --no-extras --cpu-seconds 2 ) > gpurun_out/${T}_bench150.json 2> gpurun_out/${T}_bench150.err; echo "bench150 rc=$?"
LIMU_LIB=lidar-imu-slam_b200/build/liblimu_phase.so LIMU_SPECULATE=0 timeout 300 python tools/frame_phase_timing.py > gpurun_out/${T}_phase.txt 2>&1
echo "phase rc=$?"; cat gpurun_out/${T}_phase.txt
grep -h '^{' gpurun_out/${T}_bench20.json gpurun_out/${T}_bench150.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['steps'], round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), d['e2e'].get('windows_scans_per_s'), 'it/scan', round(d['iterations_per_scan'],2), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'))
"

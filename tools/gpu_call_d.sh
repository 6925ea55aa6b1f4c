#!/usr/bin/env bash
# Round-2 GPU call D (one B200): block-header layout + one-round-trip eight-lane lookup in both shapes. GPU suite, bench (kernel-mode record), ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-d}"
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err
echo "bench20 rc=$?"
LIMU_LIB=lidar-imu-slam_b200/build/liblimu_phase.so LIMU_SPECULATE=0 timeout 300 python tools/frame_phase_timing.py > gpurun_out/${T}_phase.txt 2>&1
echo "phase rc=$?"; cat gpurun_out/${T}_phase.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_voxelize|k_icp|k_frame' -c 120 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 5 --no-extras --repeats 1 --cpu-seconds 1 > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc=$?"
grep -h '^{' gpurun_out/${T}_bench20.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), 'it/scan', round(d['iterations_per_scan'],2), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'), (d.get('parity') or {}).get('max_dt_m'))
    print(json.dumps(d.get('roofline_kernel_mode',{}).get('cases')))
    print('tracking', d.get('workload_tracking',{}).get('value'), d.get('workload_tracking',{}).get('parity'))
"

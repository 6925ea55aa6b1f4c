#!/usr/bin/env bash
# GPU call V7 (one B200): grid-wide IQR ranking for up to 16384 candidates -- GPU suite, phase timing of c3, default bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-v7}"
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/${T}_pytest.log | cut -c1-300
LIMU_PT_BG=1 LIMU_PT_WORKLOAD=c3 LIMU_LIB=lidar-imu-slam_b200/build/liblimu_phase.so LIMU_SPECULATE=0 timeout 400 python tools/frame_phase_timing.py > gpurun_out/${T}_pt_c3.txt 2>&1; echo "phase rc=$?"; head -3 gpurun_out/${T}_pt_c3.txt | cut -c1-600
( time timeout 1200 python bench.py ) > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench default rc=$?"; tail -3 gpurun_out/${T}_bench_default.err
grep -h '^{' gpurun_out/${T}_bench_default.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['steps'], round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), 'it/scan', round(d['iterations_per_scan'],2), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'), 'launches', d['gpu_launches'], 'roofline', round(d['roofline']['frac'],4), d['roofline']['share_of_step'])
    print('  km', json.dumps([(c['queries'], c['us_per_iter'], c['frac']) for c in d.get('roofline_kernel_mode',{}).get('cases',[])]), 'tracking', d.get('workload_tracking',{}).get('value'), d.get('workload_tracking',{}).get('parity',{}).get('ok'), 'mode3', d.get('icp_mode_3',{}).get('value'), 'loop', d.get('loop_closure_regime',{}).get('value'))
    c3 = d.get('workload_c3', {})
    print('  c3', c3.get('value'), c3.get('windows_scans_per_s'), c3.get('stage_ms_per_step'), (c3.get('parity') or {}).get('ok'), c3.get('error'))
"

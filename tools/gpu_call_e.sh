#!/usr/bin/env bash
# Round-2 GPU call E (one B200): classic loop with the solver warp (overlapped tail), cooperative bandwidth pass v2, occupancy variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-e}"
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err
echo "bench20 rc=$?"
for v in bw2 bw4; do
  LIMU_LIB=lidar-imu-slam_b200/build/liblimu_$v.so timeout 300 python tools/kernel_mode_bench.py > gpurun_out/${T}_km_$v.json 2> gpurun_out/${T}_km_$v.err
  echo "km $v rc=$?"
done
LIMU_LIB=lidar-imu-slam_b200/build/liblimu_phase.so LIMU_SPECULATE=0 timeout 300 python tools/frame_phase_timing.py > gpurun_out/${T}_phase.txt 2>&1
echo "phase rc=$?"; cat gpurun_out/${T}_phase.txt
grep -h '^{' gpurun_out/${T}_bench20.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), 'it/scan', round(d['iterations_per_scan'],2), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'), (d.get('parity') or {}).get('max_dt_m'))
    print('default(bw3)', json.dumps([(c['queries'], c['us_per_iter'], c['frac']) for c in d.get('roofline_kernel_mode',{}).get('cases',[])]))
    print('tracking', d.get('workload_tracking',{}).get('value'), (d.get('workload_tracking',{}).get('parity') or {}).get('ok'), 'mode3', d.get('icp_mode_3',{}).get('value'), 'loop', d.get('loop_closure_regime',{}).get('value'))
"
for v in bw2 bw4; do python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/${T}_km_$v.json') if l.startswith('{')][-1]); print('$v', [(c['queries'], c['us_per_iter'], c['frac']) for c in d['cases']])" ; done

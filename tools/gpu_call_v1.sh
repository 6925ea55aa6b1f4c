#!/usr/bin/env bash
# GPU call V1 (one B200): validation of HEAD after the session restart: GPU suite, smoke, default bench (all records), driver-like bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-v1}"
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/${T}_pytest.log
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${T}_smoke.log
( time timeout 1200 python bench.py ) > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench default rc=$?"; tail -3 gpurun_out/${T}_bench_default.err
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 2 ) > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err; echo "bench20 rc=$?"
grep -h '^{' gpurun_out/${T}_bench_default.json gpurun_out/${T}_bench20.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['steps'], round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), 'it/scan', round(d['iterations_per_scan'],2), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'), 'launches', d['gpu_launches'], 'roofline', round(d['roofline']['frac'],4))
    print('  km', json.dumps([(c['queries'], c['us_per_iter'], c['frac']) for c in d.get('roofline_kernel_mode',{}).get('cases',[])]), 'tracking', d.get('workload_tracking',{}).get('value'), 'mode3', d.get('icp_mode_3',{}).get('value'), 'loop', d.get('loop_closure_regime',{}).get('value'), 'cloud', json.dumps(d.get('e2e_cloud',{}))[:600], 'cpu', d.get('cpu_baseline',{}).get('value'))
"

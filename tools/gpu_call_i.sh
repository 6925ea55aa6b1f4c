#!/usr/bin/env bash
# GPU call I (one B200): default bench (all records), then ONE ncu --set full capture of the bandwidth-shape kernel in kernel mode (after the same command ran plainly).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-i}"
( time timeout 900 python bench.py ) > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench150 rc=$?"
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err; echo "bench20 rc=$?"
timeout 300 python tools/kernel_mode_bench.py --queries 4194304 --iters 3 > gpurun_out/${T}_km_plain.json 2> gpurun_out/${T}_km_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_icp_persistent -c 2 -o gpurun_out/${T}_km_full python tools/kernel_mode_bench.py --queries 4194304 --iters 3 > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc=$?"
grep -h '^{' gpurun_out/${T}_bench_default.json gpurun_out/${T}_bench20.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['steps'], round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), 'it/scan', round(d['iterations_per_scan'],2), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'))
    print('  km', json.dumps([(c['queries'], c['us_per_iter'], c['frac']) for c in d.get('roofline_kernel_mode',{}).get('cases',[])]), 'tracking', d.get('workload_tracking',{}).get('value'), 'mode3', d.get('icp_mode_3',{}).get('value'), 'loop', d.get('loop_closure_regime',{}).get('value'), 'cloud', d.get('e2e_cloud',{}).get('value'))
"

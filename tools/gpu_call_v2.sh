#!/usr/bin/env bash
# GPU call V2 (one B200): configs[2] in pipeline mode -- the new parity test and the default bench with the workload_c3 record.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-v2}"
( time timeout 600 python -m pytest tests/test_bench_parity.py -m gpu -q -x -k c3 ) > gpurun_out/${T}_pytest_c3.log 2>&1
echo "pytest c3 rc=$?"; tail -12 gpurun_out/${T}_pytest_c3.log
( time timeout 1200 python bench.py ) > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench default rc=$?"; tail -3 gpurun_out/${T}_bench_default.err
grep -h '^{' gpurun_out/${T}_bench_default.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['steps'], round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), 'it/scan', round(d['iterations_per_scan'],2), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'), 'launches', d['gpu_launches'], 'roofline', round(d['roofline']['frac'],4), d['roofline']['share_of_step'])
    print('  c3', json.dumps(d.get('workload_c3')))
"

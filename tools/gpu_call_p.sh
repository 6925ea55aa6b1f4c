#!/usr/bin/env bash
# GPU call P (one B200): pipelined odometry path (gated k_voxelize beside the map update, loop launched ahead). GPU suite, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-p}"
( time timeout 600 python -m pytest tests/test_speculate.py tests/test_bench_parity.py -m gpu -x -q ) > gpurun_out/${T}_pytest_spec.log 2>&1
echo "pytest spec rc=$?"; tail -15 gpurun_out/${T}_pytest_spec.log
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/${T}_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 --no-extras ) > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err; echo "bench20 rc=$?"
( time timeout 900 python bench.py --steps 150 --warmup 5 --no-extras --cpu-seconds 2 ) > gpurun_out/${T}_bench150.json 2> gpurun_out/${T}_bench150.err; echo "bench150 rc=$?"
( time timeout 900 python bench.py --steps 150 --warmup 5 --no-extras --cpu-seconds 2 --no-speculate ) > gpurun_out/${T}_bench150_plain.json 2> gpurun_out/${T}_bench150_plain.err; echo "bench150 plain rc=$?"
tail -3 gpurun_out/${T}_bench20.err
grep -h '^{' gpurun_out/${T}_bench20.json gpurun_out/${T}_bench150.json gpurun_out/${T}_bench150_plain.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['steps'], round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), d['e2e'].get('windows_scans_per_s'), 'it/scan', round(d['iterations_per_scan'],2), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}))
"

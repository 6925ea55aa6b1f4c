#!/usr/bin/env bash
# GPU call STAB (one B200): the GPU suite three times in a row (the pipelined path has timing-dependent interleavings), then the bench's icp_mode 3 record
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/stab_pytest_$i.log 2>&1; echo "run $i rc=$? $(tail -1 gpurun_out/stab_pytest_$i.log)"
done
timeout 900 python bench.py --steps 40 --warmup 5 --icp-mode 3 --no-extras --cpu-seconds 1 > gpurun_out/stab_mode3.json 2> gpurun_out/stab_mode3.err; echo "mode3 rc=$?"
grep -h '^{' gpurun_out/stab_mode3.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('icp_mode 3:', round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), 'it/scan', round(d['iterations_per_scan'],2), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'))
"

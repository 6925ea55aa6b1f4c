#!/usr/bin/env bash
# GPU call AE (one B200): experiment -- one shared-memory carve-out for all kernels of the pipelined chain
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in none 28 none 28 50 100; do
  if [ "$v" = "none" ]; then unset LIMU_EXP_CARVEOUT; else export LIMU_EXP_CARVEOUT=$v; fi
  timeout 600 python bench.py --steps 150 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/ae_bench_$v.json 2> gpurun_out/ae_bench_$v.err
  grep -h '^{' gpurun_out/ae_bench_$v.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('carveout $v', round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), 'parity', (d.get('parity') or {}).get('ok'))
"
done

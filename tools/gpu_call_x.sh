#!/usr/bin/env bash
# GPU call X (one B200): experiment -- footprint of the voxel table (TLB / L2 locality of the lookups)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfgv in "default" "131072 30000" "65536 30000" "40000 20000"; do
  set -- $cfgv
  if [ "$1" = "default" ]; then unset LIMU_EXP_MAP_VOXELS LIMU_EXP_INCOMING; else export LIMU_EXP_MAP_VOXELS=$1 LIMU_EXP_INCOMING=$2; fi
  timeout 600 python bench.py --steps 150 --warmup 5 --no-extras --cpu-seconds 1 > gpurun_out/x_bench_$1.json 2> gpurun_out/x_bench_$1.err
  grep -h '^{' gpurun_out/x_bench_$1.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$cfgv', round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), 'it/scan', round(d['iterations_per_scan'],2), d['stage_ms_per_step'], 'voxels', d.get('map_voxels'), 'parity', (d.get('parity') or {}).get('ok'))
"
done

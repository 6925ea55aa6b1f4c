#!/usr/bin/env bash
# Round-2 GPU call F (one B200): staged (cp.async -> shared memory) pass of the bandwidth shape; occupancy variants.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-f}"
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
timeout 300 python tools/kernel_mode_bench.py > gpurun_out/${T}_km_bw2.json 2> gpurun_out/${T}_km_bw2.err; echo "km bw2(default) rc=$?"
for v in bw3 bw4; do
  LIMU_LIB=lidar-imu-slam_b200/build/liblimu_$v.so timeout 300 python tools/kernel_mode_bench.py > gpurun_out/${T}_km_$v.json 2> gpurun_out/${T}_km_$v.err
  echo "km $v rc=$?"
done
timeout 300 python tools/kernel_mode_bench.py --cap 10 --voxel 1.0 --fill 10 > gpurun_out/${T}_km_cap10.json 2> gpurun_out/${T}_km_cap10.err; echo "km cap10 rc=$?"
for v in bw2 bw3 bw4 cap10; do python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/${T}_km_$v.json') if l.startswith('{')][-1]); print('$v', [(c['queries'], c['us_per_iter'], c['frac'], c['k_bar']) for c in d['cases']])" ; done

#!/usr/bin/env bash
# Round-2 GPU call C (one B200): the cluster latency shape. GPU suite, bench A/B (LIMU_CLUSTER_LOOP), phase timing of both shapes, ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-c}"
( time timeout 600 python -m pytest tests/test_speculate.py -m gpu -x -q ) > gpurun_out/${T}_pytest_first.log 2>&1
echo "pytest(first) rc=$?"; tail -4 gpurun_out/${T}_pytest_first.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
( time timeout 600 python bench.py --steps 20 --warmup 5 --no-extras ) > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err
echo "bench20 rc=$?"
( time LIMU_CLUSTER_LOOP=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 2 ) > gpurun_out/${T}_bench20_classic.json 2> gpurun_out/${T}_bench20_classic.err
echo "bench20 classic rc=$?"
( time timeout 900 python bench.py ) > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
echo "bench150 rc=$?"
LIMU_LIB=lidar-imu-slam_b200/build/liblimu_phase.so LIMU_SPECULATE=0 timeout 300 python tools/frame_phase_timing.py > gpurun_out/${T}_phase_cluster.txt 2>&1
echo "phase cluster rc=$?"
LIMU_LIB=lidar-imu-slam_b200/build/liblimu_phase.so LIMU_SPECULATE=0 LIMU_CLUSTER_LOOP=0 timeout 300 python tools/frame_phase_timing.py > gpurun_out/${T}_phase_classic.txt 2>&1
echo "phase classic rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_voxelize|k_icp|k_frame' -c 120 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 20 --warmup 5 --no-extras --repeats 1 --cpu-seconds 1 > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc=$?"
cat gpurun_out/${T}_phase_cluster.txt gpurun_out/${T}_phase_classic.txt
grep -h '^{' gpurun_out/${T}_bench20.json gpurun_out/${T}_bench20_classic.json gpurun_out/${T}_bench_default.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), 'it/scan', round(d['iterations_per_scan'],2), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'), (d.get('parity') or {}).get('max_dt_m'))
"

#!/usr/bin/env bash
# Round-2 GPU call A (one B200): full GPU suite, driver-like bench, default bench, speculation A/B, phase timing, bulk-gather micro-benchmark,
# ncu launch list. Every step logs to gpurun_out/ and none blocks the next.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
( time timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/a_bench_driverlike.json 2> gpurun_out/a_bench_driverlike.err
echo "bench20 rc=$?"
( time timeout 600 python bench.py --steps 20 --warmup 5 --no-speculate --no-extras ) > gpurun_out/a_bench_nospec.json 2> gpurun_out/a_bench_nospec.err
echo "bench20 nospec rc=$?"
( time timeout 900 python bench.py ) > gpurun_out/a_bench_default.json 2> gpurun_out/a_bench_default.err
echo "bench150 rc=$?"
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err
echo "ref rc=$?"
LIMU_LIB=lidar-imu-slam_b200/build/liblimu_phase.so LIMU_SPECULATE=0 timeout 300 python tools/frame_phase_timing.py > gpurun_out/a_phase.txt 2>&1
echo "phase rc=$?"
timeout 300 tools/experimental/build/bulk_gather > gpurun_out/a_bulk_gather.txt 2>&1
echo "bulk rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 20 --warmup 5 --no-extras --repeats 1 --cpu-seconds 1 > gpurun_out/a_ncu.log 2>&1
echo "ncu rc=$?"
head -c 1500 gpurun_out/a_bench_driverlike.json

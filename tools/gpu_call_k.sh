#!/usr/bin/env bash
# GPU call K (one B200): ncu --set full of the final library's kernels at configs[2] size -- kernel mode (4 M queries, 3 iterations per launch)
# and pipeline mode (512 000 points/scan, ~45 M-point map) -- each only after the same command ran plainly with exit 0.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-k}"
KM="python tools/kernel_mode_bench.py --queries 4194304 --iters 3"
timeout 300 $KM > gpurun_out/${T}_km_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_icp_persistent' --launch-skip 2 -c 2 -o gpurun_out/${T}_km_full $KM > gpurun_out/${T}_km_ncu.log 2>&1
echo "kernel mode ncu rc=$?"; tail -2 gpurun_out/${T}_km_plain.log | cut -c1-400
C3="python tools/c3_gpu_only.py"
timeout 300 $C3 > gpurun_out/${T}_c3_plain.log 2>&1 && \
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_icp_persistent|k_voxelize_lean|k_frame_update' --launch-skip 15 -c 6 -o gpurun_out/${T}_c3_full $C3 > gpurun_out/${T}_c3_ncu.log 2>&1
echo "c3 ncu rc=$?"; tail -1 gpurun_out/${T}_c3_plain.log | cut -c1-600
ls -la gpurun_out/${T}_*.ncu-rep

#!/usr/bin/env bash
# GPU call F2 (one B200): ONE ncu --set full capture of the pipelined path's kernels (after the same command ran plainly with exit 0).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-f2}"
CMD="python bench.py --steps 12 --warmup 5 --no-extras --cpu-seconds 1 --repeats 1"
timeout 600 $CMD > gpurun_out/${T}_plain.json 2> gpurun_out/${T}_plain.err && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_icp_persistent|k_voxelize_lean|k_frame_update' --launch-skip 30 -c 9 -o gpurun_out/${T}_frame_full $CMD > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 gpurun_out/${T}_ncu.log
python -c "
import __graft_entry__ as g
print('source_hash', g.load_package().source_hash())
" | tee gpurun_out/${T}_hash.txt

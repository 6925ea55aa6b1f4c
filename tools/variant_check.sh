#!/usr/bin/env bash
# Build a compile-time VARIANT of liblimu_cuda (extra -D flags) next to the default library and check it on the GPU box:
# the whole GPU suite through LIMU_LIB, then bench.py with both libraries.
#   tools/variant_check.sh spec -DLIMU_SPECULATIVE_VOXELIZE          (build here, where nvcc cross-compiles)
#   gpurun -- 'tools/variant_check.sh --run spec'                     (run on the B200 box; the .so travels with the snapshot)
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
PKG="$ROOT/lidar-imu-slam_b200"
if [[ "${1:-}" != "--run" ]]; then
    NAME="$1"; shift
    mkdir -p "$PKG/build"
    nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off "$@" \
         -shared -o "$PKG/build/liblimu_$NAME.so" "$PKG"/csrc/*.cu -lcudart_static -lpthread -ldl -lrt
    echo "built $PKG/build/liblimu_$NAME.so with $*"
    exit 0
fi
NAME="$2"
LIBV="$PKG/build/liblimu_$NAME.so"
mkdir -p "$ROOT/gpurun_out"
cd "$ROOT"
# test_library_loaded_is_in_tree and the C++ consumer pin the default library by design
LIMU_LIB="$LIBV" python -m pytest tests -m gpu -q -k "not library_loaded and not cpp_consumer" 2>&1 | tail -5
python bench.py --steps 150 --warmup 5 --cpu-seconds 2 > "gpurun_out/bench_default.json" 2> "gpurun_out/bench_default.err"
LIMU_LIB="$LIBV" python bench.py --steps 150 --warmup 5 --cpu-seconds 2 > "gpurun_out/bench_$NAME.json" 2> "gpurun_out/bench_$NAME.err"
python - "$NAME" <<'PY'
import json, sys
a = json.load(open("gpurun_out/bench_default.json")); b = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
for k in ("value", "ms_per_step"):
    print(k, "default", round(a[k], 4), sys.argv[1], round(b[k], 4))
print("e2e default", round(a["e2e"]["value"], 1), sys.argv[1], round(b["e2e"]["value"], 1))
print("last_pose equal to 1e-6:", max(abs(x - y) for x, y in zip(a["last_pose"], b["last_pose"])) < 1e-6)
PY

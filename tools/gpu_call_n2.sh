#!/usr/bin/env bash
# Two-GPU call: the multi-GPU test, then the bench as the driver launches it at N = 2 (replicas + the point-sharded record).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-n2}"
( time timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q ) > gpurun_out/${T}_pytest_mg.log 2>&1; echo "pytest multi_gpu rc=$?"; tail -3 gpurun_out/${T}_pytest_mg.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "bench n2 rc=$?"
grep -h '^{' gpurun_out/${T}_bench_n2.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['n_gpus'], round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), d.get('per_rank'), 'sharded', d.get('sharded'))
"

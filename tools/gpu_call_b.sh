#!/usr/bin/env bash
# Round-2 GPU call B (N B200s of one box, N = $1, default 2): the multi-GPU test, bench.py under torchrun exactly as the driver launches it,
# and the sharded ICP check incl. the un-fused NCCL baseline.
N="${1:-2}"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/b${N}_topo.txt 2>&1
( time timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q ) > gpurun_out/b${N}_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/b${N}_pytest.log
NS="1 $N"; [ "$N" = "8" ] && NS="1 4 8"
for n in $NS; do
  if [ "$n" = "1" ]; then
    ( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extras ) > gpurun_out/b${N}_bench_n1.json 2> gpurun_out/b${N}_bench_n1.err
  else
    ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 5 ) > gpurun_out/b${N}_bench_n$n.json 2> gpurun_out/b${N}_bench_n$n.err
  fi
  echo "bench n=$n rc=$?"
done
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 5 --warmup 2 ) > gpurun_out/b${N}_bench_ref.json 2> gpurun_out/b${N}_bench_ref.err
echo "ref rc=$?"
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tools/sharded_icp_check.py --queries 100000,4194304 --voxels 2.5e6 --iters 20 --reps 3 --alarm 500 ) > gpurun_out/b${N}_sharded.jsonl 2> gpurun_out/b${N}_sharded.err
echo "sharded rc=$?"
cat gpurun_out/b${N}_sharded.jsonl
grep -h '^{' gpurun_out/b${N}_bench_n*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['n_gpus'], round(d['value'],1), d['windows_scans_per_s'], d['per_rank'], json.dumps(d.get('sharded'))[:600])
"

#!/usr/bin/env bash
# GPU call H (one B200): staged pass with four lanes per block. GPU suite + kernel-mode record (cap 20 and cap 10 maps).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-h}"
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/${T}_pytest.log
timeout 300 python tools/kernel_mode_bench.py > gpurun_out/${T}_km.json 2> gpurun_out/${T}_km.err; echo "km rc=$?"
timeout 300 python tools/kernel_mode_bench.py --cap 10 --voxel 1.0 --fill 10 > gpurun_out/${T}_km_cap10.json 2> gpurun_out/${T}_km_cap10.err; echo "km cap10 rc=$?"
for v in km km_cap10; do python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/${T}_$v.json') if l.startswith('{')][-1]); print('$v', [(c['queries'], c['us_per_iter'], c['frac'], c['frac_with_source_writeback'], c['k_bar']) for c in d['cases']])" ; done

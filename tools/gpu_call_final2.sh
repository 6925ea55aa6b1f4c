#!/usr/bin/env bash
# Final evidence call, second edition (one B200): GPU suite and the 240-seed random campaign first -- the call stops there if either fails --
# then smoke, default bench (all records), driver-like bench, reference arm, and (each only after the same command ran plainly with exit 0)
# the ncu launch list and ONE ncu --set full capture of the pipelined path's kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="${1:-fin2}"
BUDGET="${2:-700}"   # seconds this script may use in all: every leg is cut to what is left of it
left() { local l=$(( BUDGET - SECONDS )); [ $l -lt 5 ] && l=5; [ $l -gt $1 ] && l=$1; echo $l; }
python -c "
import __graft_entry__ as g
print('source_hash', g.load_package().source_hash())
" | tee gpurun_out/${T}_hash.txt
( time timeout $(left 600) python -m pytest ${PYT:-tests} -m gpu -q ) > gpurun_out/${T}_pytest.log 2>&1; RC1=$?
echo "pytest rc=$RC1"; tail -4 gpurun_out/${T}_pytest.log | cut -c1-300
RC2=0
[ "${CAMPAIGN:-1}" = 1 ] && { ( time LIMU_RANDOM_SEEDS=16-256 timeout $(left 400) python -m pytest tests/test_speculate.py -m gpu -q -k random -n 4 -p no:cacheprovider ) > gpurun_out/${T}_campaign.log 2>&1; RC2=$?; }
echo "campaign rc=$RC2"; [ "${CAMPAIGN:-1}" = 1 ] && tail -12 gpurun_out/${T}_campaign.log | cut -c1-300
if [ $RC1 -ne 0 ] || [ $RC2 -ne 0 ]; then echo "STOP: tests failed"; exit 1; fi
( time timeout $(left 300) python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke ok" gpurun_out/${T}_smoke.log
CMD="python bench.py --steps 12 --warmup 5 --no-extras --cpu-seconds 1 --repeats 1"
timeout $(left 400) $CMD > gpurun_out/${T}_plain_for_ncu.json 2> gpurun_out/${T}_plain_for_ncu.err && \
timeout $(left 600) ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_voxelize|k_icp|k_frame|k_gate' -c 400 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu_list.log 2>&1
echo "ncu launch list rc=$?"
timeout $(left 900) ncu --set full --clock-control none --import-source on -k regex:'k_icp_persistent|k_voxelize_lean|k_frame_update' --launch-skip 30 -c 9 -o gpurun_out/${T}_frame_full $CMD > gpurun_out/${T}_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/${T}_ncu_full.log
( time timeout $(left 600) python bench.py --steps 20 --warmup 5 ) > gpurun_out/${T}_bench20.json 2> gpurun_out/${T}_bench20.err; echo "bench20 rc=$?"
( time timeout $(left 400) python bench.py --impl reference --steps 20 --warmup 5 ) > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "bench ref rc=$?"
( time timeout $(left 900) python bench.py ) > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench default rc=$?"; tail -3 gpurun_out/${T}_bench_default.err
grep -h '^{' gpurun_out/${T}_bench_default.json gpurun_out/${T}_bench20.json gpurun_out/${T}_bench_ref.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d.get('impl') == 'reference':
        print('reference arm', d.get('value'), d.get('unit'), d.get('cpu_baseline'))
        continue
    print(d['steps'], round(d['value'],1), d['windows_scans_per_s'], 'e2e', round(d['e2e']['value'],1), 'it/scan', round(d['iterations_per_scan'],2), d['stage_ms_per_step'], 'parity', (d.get('parity') or {}).get('ok'), 'launches', d['gpu_launches'], 'roofline', round(d['roofline']['frac'],4))
    c3 = d.get('workload_c3', {})
    print('  km', json.dumps([(c['queries'], c['us_per_iter'], c['frac']) for c in d.get('roofline_kernel_mode',{}).get('cases',[])]), 'tracking', d.get('workload_tracking',{}).get('value'), 'mode3', d.get('icp_mode_3',{}).get('value'), d.get('icp_mode_3',{}).get('iterations_per_scan'), d.get('icp_mode_3',{}).get('last_pose_xy'), d.get('icp_mode_3',{}).get('true_xy'), 'loop', d.get('loop_closure_regime',{}).get('value'), 'cloud', d.get('e2e_cloud',{}).get('value'), (d.get('e2e_cloud',{}).get('with_prefetch') or {}).get('value'), 'cpu', d.get('cpu_baseline',{}).get('value'), 'c3', c3.get('value'), (c3.get('parity') or {}).get('ok'))
"

"""Turn the artefacts of tools/gpu_call_final.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/:
r2_final_bench.json, r2_launch_list.json (+ r2_launches.csv), r2_frame_kernels_ncu.json.   python tools/collect_final.py [tag]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = sys.argv[1] if len(sys.argv) > 1 else "fin"
G = lambda name: os.path.join(ROOT, "gpurun_out", f"{T}_{name}")
P = lambda name: os.path.join(ROOT, "profiles", name)


def jline(path):
    for l in open(path):
        if l.startswith("{"):
            return json.loads(l)
    return None


src_hash = open(G("hash.txt")).read().split()[-1]
pytest_tail = open(G("pytest.log")).read().strip().splitlines()[-5:]
out = {"what": f"final GPU call of round 2 (one B200, library source hash {src_hash}): GPU suite, smoke(), python bench.py (default: K=150, W=5, all records), "
               "python bench.py --steps 20 --warmup 5, python bench.py --impl reference --steps 20 --warmup 5",
       "gpu_suite": [l for l in pytest_tail if "passed" in l or "failed" in l],
       "smoke": [l for l in open(G("smoke.log")).read().splitlines() if "smoke ok" in l],
       "default": jline(G("bench_default.json")), "k20": jline(G("bench20.json")), "reference_arm_k20": jline(G("bench_ref.json"))}
json.dump(out, open(P("r2_final_bench.json"), "w"), indent=1)

# launch list
rows = [r for r in csv.reader(l for l in open(G("launches.csv")) if l.startswith('"'))]
hdr, body = rows[0], rows[1:]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = {}
for r in body:
    name = r[kn].split("(")[0]
    agg.setdefault(name, []).append(float(r[mv].replace(",", "")) / 1e3)
tot = sum(sum(v) for v in agg.values())
kernels = sorted(({"kernel": k, "launches": len(v), "mean_us": round(sum(v) / len(v), 2), "share_of_listed_time": round(sum(v) / tot, 4)} for k, v in agg.items()),
                 key=lambda d: -d["share_of_listed_time"])
open(P("r2_launches.csv"), "w").write(open(G("launches.csv")).read())
json.dump({"what": "ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_voxelize|k_icp|k_frame|k_gate' -c 400 --csv python bench.py --steps 12 --warmup 5 "
                   "--no-extras --cpu-seconds 1 --repeats 1 (one B200, after the same command ran plainly with exit 0). Serialised, cold cache: shares, not absolutes.",
           "source_hash": src_hash, "kernels": kernels,
           "note": "bench.py's stage events measure the kernels as they run in the pipelined path, where k_voxelize_lean and k_frame_update run BESIDE each other (and slow "
                   "each other down); ncu serialises every launch, so the loop kernel's share of the listed time is larger than its share of the sum of the stage times "
                   "and smaller than its share of the step period."}, open(P("r2_launch_list.json"), "w"), indent=1)
print(json.dumps(kernels, indent=1))

# ncu --set full
rep = G("frame_full.ncu-rep")
if os.path.exists(rep):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ncu_summary
    tmp = P("r2_frame_kernels_ncu.json")
    ncu_summary.main(rep, tmp)
    d = json.load(open(tmp))
    d = {"what": "ncu --set full --clock-control none --import-source on -k regex:'k_icp_persistent|k_voxelize_lean|k_frame_update' --launch-skip 30 -c 9 python bench.py "
                 "--steps 12 --warmup 5 --no-extras --cpu-seconds 1 --repeats 1 (one B200; the same command ran plainly first, exit 0). Per-launch numbers are cold-cache "
                 "and serialised (ncu replays every kernel ~39 times).",
         "source_hash": src_hash, "launches": d["launches"]}
    json.dump(d, open(tmp, "w"), indent=1)
    print("ncu summary:", len(d["launches"]), "launches")

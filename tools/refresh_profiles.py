#!/usr/bin/env python3
"""Copy the round's measurement artefacts from gpurun_out/ into profiles/ and rebuild the derived summaries:
launch-list table (profiles/<tag>_launches.md), full-set ncu summary (profiles/<tag>_ncu_summary.md) and the DRAM traffic
per launch that bench.py reports as roofline.traffic (profiles/traffic.json).  Usage: python tools/refresh_profiles.py [tag]"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for f in ("%s_bench.json", "%s_bench_reference_arm.json", "%s_launches.csv", "%s_bench_2gpu.json", "%s_bench_8gpu.json", "%s_path_sweep.json",
          "%s_codec_timings.json", "%s_agg_timings.json", "%s_latency.json", "%s_fp64_peaks.json", "%s_msm_phases.json"):
    src = os.path.join(G, f % tag)
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, f % tag))
rows = list(csv.reader(open(os.path.join(G, "%s_launches.csv" % tag))))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hi]; ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "").replace("b381::", "")
    tot[name] += float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0); cnt[name] += 1
T = sum(tot.values())
out = ["# ncu launch list of `python bench.py --steps 2 --warmup 3 --cpu-seconds 1` (gpu__time_duration.sum, --clock-control none; "
       "cold-cache, serialised: compare shares), build %s" % tag, "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
out += ["| %s | %d | %.3f | %.1f %% |" % (k, cnt[k], v, 100 * v / T) for k, v in tot.most_common()]
open(os.path.join(P, "%s_launches.md" % tag), "w").write("\n".join(out) + "\n")
rep = os.path.join(G, "%s_pairing.ncu-rep" % tag)
if os.path.exists(rep):
    summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
    head = ("# ncu --set full summaries, build %s (thread-per-pairing kernels, 2^16 pairings per launch)\n\n"
            "Command (gpurun, 1 x B200): `B381_PATH=thread ncu --set full --clock-control none --import-source on -k regex:\"k_final_exp|k_miller_loop\" -s 2 -c 2 "
            "python tools/profile_run.py --reps 2`. Launch list of the bench command: `profiles/%s_launches.md` / `.csv`; bench line: "
            "`profiles/%s_bench.json`; experiments: `profiles/%s_experiments.md`.\n\n" % (tag, tag, tag, tag))
    open(os.path.join(P, "%s_ncu_summary.md" % tag), "w").write(head + summ)
    tr = json.loads(subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "--traffic", rep], capture_output=True, text=True).stdout)
    tr["source"] = "profiles/%s_ncu_summary.md (ncu --set full, 2^16 pairings per launch)" % tag
    sys.path.insert(0, ROOT)
    import bench
    tr["source_stamp"] = bench.pairing_source_hash()       # bench.py reports the traffic only while the sources still hash to this
    tr["commit"] = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    json.dump(tr, open(os.path.join(P, "traffic.json"), "w"))
# further full-set captures of the round: lane kernels, MSM chunk sums
for name, title in (("lanes", "two-lane / four-lane pairing kernels at 2^16 pairings"), ("msm", "k_msm_chunk_sum at 2^22 points")):
    rep2 = os.path.join(G, "%s_%s.ncu-rep" % (tag, name))
    if os.path.exists(rep2):
        summ = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep2], capture_output=True, text=True).stdout
        open(os.path.join(P, "%s_ncu_summary_%s.md" % (tag, name)), "w").write("# ncu --set full summary, build %s: %s\n\n" % (tag, title) + summ)
d = json.load(open(os.path.join(P, "%s_bench.json" % tag)))
print("value %.0f e2e %.0f frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
print(json.dumps(d["aggregate"]))

"""per-kernel device times + stream gaps of one config step (bench.py's Rig, CUDA events around every launch)"""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import bench
from tools import benchkit as bk
name = sys.argv[1]
rows = int(sys.argv[2])
rig = bench.Rig(torch, None, 0, 1, 0)
w = {"c3": bk.c3_workload, "c4": bk.c4_workload}[name](rows)
for (r, c) in w.referenced():
    rig.upload_fn(r, c, w.rows(r), lambda b, n, r=r, c=c: w.column_t(torch, r, c, b, n))
text = w.text(0)
for _ in range(3):
    out = rig.run(text)
ms, launches, out, wall = rig.timed(text, 5, 0)
rig.eng.profile(True)
t0 = time.perf_counter()
for _ in range(5):
    rig.run(text)
wall_prof = (time.perf_counter() - t0) / 5 * 1e3
prof = rig.eng.profile_read()
rig.eng.profile(False)
k = {a: round(b["ms"] / 5, 4) for a, b in prof.items() if not a.startswith("gap_")}
g = {a: round(b["ms"] / 5, 4) for a, b in prof.items() if a.startswith("gap_") and b["ms"] / 5 > 0.02}
print(json.dumps({"config": name, "rows": rows, "ms_per_step": ms, "wall": wall, "launches": launches, "wall_with_events": wall_prof,
                  "kernel_ms_sum": sum(k.values()), "gap_ms_sum": sum(b["ms"] for a, b in prof.items() if a.startswith("gap_")) / 5,
                  "kernels": dict(sorted(k.items(), key=lambda kv: -kv[1])), "gaps": dict(sorted(g.items(), key=lambda kv: -kv[1])[:25]),
                  "launch_counts": {a: b["launches"] // 5 for a, b in prof.items() if not a.startswith("gap_")}}, indent=1))

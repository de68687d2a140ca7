"""per-kernel device time of the full config-5 batch (main context = the heavy queries; worker contexts are not timed)"""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import torch
import bench
from tools import benchkit as bk
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
rig = bench.Rig(torch, None, 0, 1, 0)
w = bk.c5_workload(scale=scale, nqueries=1000)
for (r, c) in w.referenced():
    rig.upload_fn(30 + r, c, w.rows(r), lambda b, n, r=r, c=c: w.column_t(torch, r, c, b, n))
torch.cuda.empty_cache()
text = w.text(30)
rig.run(text)
t0 = time.perf_counter(); rig.run(text); wall = time.perf_counter() - t0
rig.eng.profile(True)
t0 = time.perf_counter(); rig.run(text); wall_prof = time.perf_counter() - t0
prof = rig.eng.profile_read()
k = {a: round(b["ms"], 2) for a, b in prof.items() if not a.startswith("gap_")}
g = {a: round(b["ms"], 2) for a, b in prof.items() if a.startswith("gap_") and b["ms"] > 5}
print(json.dumps({"wall_s": wall, "wall_prof_s": wall_prof, "kernel_ms_sum": sum(k.values()),
                  "gap_ms_sum": sum(b["ms"] for a, b in prof.items() if a.startswith("gap_")),
                  "kernels": dict(sorted(k.items(), key=lambda kv: -kv[1])), "gaps": dict(sorted(g.items(), key=lambda kv: -kv[1])[:20]),
                  "launches": {a: b["launches"] for a, b in prof.items() if not a.startswith("gap_")}}, indent=1))

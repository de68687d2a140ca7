"""C1-shaped batch: many small queries (latency-bound regime).  queries/s of the
host layer vs the reference binary, plus where the time goes."""
import sys, os, time, tempfile, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, qce_b200, bench
from oracle import workload as wl, qce_oracle as orc
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 200
db = wl.gen_small_db(seed=2018, scale=scale)
e = qce_b200.Engine(); lib = bench.host_lib()
e.upload_db(db)
qs = wl.gen_queries(db, nq, seed=5, max_joins=2)
text = "".join(q + "\n" for q in qs)
import ctypes as C
buf = C.create_string_buffer(1 << 20); failed = C.c_int(0)
def run():
    n = lib.qce_host_run_batch(text.encode(), buf, 1 << 20, C.byref(failed))
    return buf.value.decode()
out = run(); run()
e.timer_reset(); t = time.perf_counter()
for _ in range(3): out = run()
ms, launches = e.timer_read(); wall = (time.perf_counter() - t) / 3
print(f"rows total {sum(len(r[0]) for r in db)}, {nq} queries: {1e3*wall:.2f} ms/batch -> {nq/wall:.0f} queries/s, launches/query {launches/3/nq:.1f}, failed {failed.value}")
e.profile(True); run(); prof = e.profile_read(); e.profile(False)
kern = {k: v for k, v in prof.items() if not k.startswith("gap_")}
gaps = {k: v for k, v in prof.items() if k.startswith("gap_")}
print("kernel ms/batch", round(sum(v["ms"] for v in kern.values()), 2), "gap ms/batch", round(sum(v["ms"] for v in gaps.values()), 2))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]: print(f"   {k:32s} {v['launches']:6d} x {v['ms']:9.3f} ms")
if False:
    paths = wl.write_db(tempfile.mkdtemp(), db)
    t = time.perf_counter(); ref, err, rc = wl.run_reference(paths, text, timeout=3000); dt = time.perf_counter() - t
    t = time.perf_counter(); wl.run_reference(paths, ""); load = time.perf_counter() - t
    print(f"reference: {dt-load:.2f} s -> {nq/(dt-load):.1f} queries/s; identical stdout: {ref == out}")

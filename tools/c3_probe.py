"""C3-shaped 4-way PK-FK chain with a self-join predicate and a filter, one GPU,
through the host layer: per-kernel breakdown."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, qce_b200, bench
from oracle import workload as wl, qce_oracle as orc
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
t = time.time(); db = wl.gen_chain_db(n, nrel=4, seed=3); print("gen", round(time.time() - t, 1), "s", flush=True)
e = qce_b200.Engine(); lib = bench.host_lib()
e.upload_db(db)
q = "0 1 2 3|0.1=0.2&0.1=1.0&1.1=2.0&2.1=3.0&0.3<900|0.3 1.3 2.3 3.3\n"
out = bench.run_query(lib, q); print("result", out.strip(), flush=True)
if n <= 2_000_000:
    print("oracle agrees:", orc.run_batch(db, q) == out)
print("truth  ", wl.truth_query(orc.parse_query(q), db).strip())
for _ in range(2): bench.run_query(lib, q)
e.timer_reset()
for _ in range(3): bench.run_query(lib, q)
ms, launches = e.timer_read()
print(f"{ms/3:.2f} ms/query, {launches/3:.0f} launches, input rows/s {4*n/(ms/3/1e3):.3e}")
e.profile(True); bench.run_query(lib, q); prof = {k: v for k, v in e.profile_read().items() if not k.startswith("gap_")}; e.profile(False)
tot = sum(v["ms"] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]: print(f"   {k:20s} {v['launches']:4d} x {v['ms']:8.3f} ms {100*v['ms']/tot:5.1f}%")

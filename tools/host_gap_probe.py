"""Per-step timing of the host layer (stability / warm-up behaviour)."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, qce_b200, bench
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
e = qce_b200.Engine(); lib = bench.host_lib()
for r, seed in enumerate((1, 2)):
    for c, col in enumerate(bench.gen_relation(n, seed, n)):
        e.upload_column(r, c, col)
q = bench.QUERY.format(thr=500000)
for prof in (False, True, False):
    e.profile(prof)
    ts = []
    for _ in range(12):
        e.timer_reset(); bench.run_query(lib, q); ms, _ = e.timer_read(); ts.append(round(ms, 2))
    e.profile(False)
    print(f"profile={prof}: per-step ms {ts}")

"""Loader throughput: write C2-sized relation files, run the driver with QCE_TIMING=1."""
import os, sys, subprocess, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from oracle import workload as wl
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
d = "/tmp/qce_load_probe"; os.makedirs(d, exist_ok=True)
t = time.time()
paths = wl.write_db(d, [bench.gen_relation(n, 1, n), bench.gen_relation(n, 2, n)])
print("wrote", sum(os.path.getsize(p) for p in paths) / 1e9, "GB in", round(time.time() - t, 1), "s")
binary = os.path.join(bench.PKG, "build", "queries")
q = bench.QUERY.format(thr=500000)
for rep in range(3):
    p = subprocess.run([binary], input=wl.stdin_text(paths, q).encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       env=dict(os.environ, QCE_TIMING="1"))
    print(p.stdout.decode().strip(), "|", p.stderr.decode().strip().splitlines()[-1])
for p in paths: os.unlink(p)

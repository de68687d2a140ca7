"""config 2 at full size, the query N times (for ncu captures: no twin, no checker, no timing)"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rig = bench.Rig(torch, None, 0, 1, 0)
for r, seed in enumerate((1, 2)):
    for c, col in enumerate(bench.gen_relation(n, seed, n)):
        rig.ck(rig.lib.qce_upload_column(r, c, col.ctypes.data, n))
q = bench.QUERY.format(thr=500000)
for _ in range(reps):
    out = rig.run(q)
print(out.strip())

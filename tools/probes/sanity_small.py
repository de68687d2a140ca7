"""small end-to-end pass over the round-2 kernels for compute-sanitizer (memcheck): bulk-copy partition and
count sort, the big-sub-bucket route, join, elided chain, checksums"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from oracle import workload as wl, qce_oracle as orc
rig = bench.Rig(torch, None, 0, 1, 0)
e = rig.eng
rng = np.random.default_rng(1)
for n, kind in ((2_200_000, "uniform"), (2_500_000, "heavy")):
    keys = rng.integers(0, 1 << 27, n, dtype=np.uint64)
    if kind == "heavy":
        keys[rng.random(n) < 0.3] = 77_777_777
        for k in range(6):
            keys[rng.choice(n, 5000 + 1500 * k, replace=False)] = 1_000_003 * (k + 1)
    t = e.tuples_from_host(keys, np.arange(n, dtype=np.uint64))
    e.sort_tuples(t)
    assert e.is_sorted(t)
    k, p = e.tuples_to_host(t)
    assert (k == np.sort(keys)).all()
    e.tuples_free(t)
    print("sort", kind, "ok", flush=True)
db = wl.gen_pair_db(1_500_000, 1_500_000)
for r, cols in enumerate(db):
    for c, a in enumerate(cols):
        rig.upload_host(r, c, a)
q = "0 1|0.1=1.1&0.2>500000|0.0 1.0 1.2\n"
assert rig.run(q) == orc.run_batch(db, q)
print("c2 ok", flush=True)
db = wl.gen_chain_db(400_000, nrel=4, seed=3)
for r, cols in enumerate(db):
    for c, a in enumerate(cols):
        rig.upload_host(10 + r, c, a)
q = "10 11 12 13|0.1=0.2&0.1=1.0&1.1=2.0&2.1=3.0&0.3<900|0.3 1.3 2.3 3.3\n"
want = orc.run_batch(db, q.replace("10 11 12 13", "0 1 2 3"))
assert rig.run(q) == want
print("c3 ok", flush=True)

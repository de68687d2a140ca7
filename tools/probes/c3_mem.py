"""arena balance over repeated config-3 steps: (reserved, in use) after every step"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import bench
from tools import benchkit as bk
rig = bench.Rig(torch, None, 0, 1, 0)
w = bk.c3_workload(62_500_000)
for (r, c) in w.referenced():
    rig.upload_fn(10 + r, c, w.rows(r), lambda b, n, r=r, c=c: w.column_t(torch, r, c, b, n))
torch.cuda.empty_cache()
text = w.text(10)
for s in range(int(os.environ.get("STEPS", "14"))):
    rig.run(text)
    rig.eng.sync()
    print(s, rig.eng.mempool_stats(), flush=True)

import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
mode = sys.argv[1]
import numpy as np
if mode in ("torch", "pinned"):
    import torch
    torch.cuda.set_device(0)
import qce_b200, bench
n = 100_000_000
e = qce_b200.Engine(0); lib = bench.host_lib()
keep = []
for r, seed in enumerate((1, 2)):
    for c, col in enumerate(bench.gen_relation(n, seed, n)):
        if mode == "pinned":
            t = torch.from_numpy(col.view(np.int64)).pin_memory(); keep.append(t)
            e.lib.qce_upload_column(r, c, t.data_ptr(), n)
        else:
            e.upload_column(r, c, col)
q = bench.QUERY.format(thr=500000)
ts = []
for _ in range(14):
    e.timer_reset(); bench.run_query(lib, q); ms, _ = e.timer_read(); ts.append(round(ms, 2))
print(mode, ts)

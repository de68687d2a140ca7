"""BASELINE config 4 on one GPU through the host layer: 3-way join with a Zipf(1.2)
foreign key (the heaviest key holds ~18 % of the rows), per-kernel breakdown.
Full-size check (size-independent properties): R0.c1 -> R1.c0 and R1.c1 -> R2.c0 are
FK -> PK joins, so the result has exactly len(R0) rows and the checksum of 0.2 is
sum(R0.c2); the scaled twin is compared with the oracle / truth in tests/."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qce_b200, bench
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(4)
t0 = time.time()
w = torch.arange(1, n + 1, dtype=torch.float64, device=dev).pow_(-1.2)
cdf = torch.cumsum(w / w.sum(), 0); del w
z = torch.searchsorted(cdf, torch.rand(n, dtype=torch.float64, device=dev, generator=g)).clamp_(0, n - 1); del cdf
perm = torch.randperm(n, device=dev, generator=g)
cols = {(0, 0): torch.arange(n, dtype=torch.int64, device=dev), (0, 1): perm[z], (0, 2): torch.randint(0, 1000, (n,), dtype=torch.int64, device=dev, generator=g),
        (1, 0): torch.randperm(n, device=dev, generator=g), (1, 1): torch.randint(0, n, (n,), dtype=torch.int64, device=dev, generator=g),
        (1, 2): torch.randint(0, 1000, (n,), dtype=torch.int64, device=dev, generator=g),
        (2, 0): torch.randperm(n, device=dev, generator=g), (2, 1): torch.randint(0, 1000, (n,), dtype=torch.int64, device=dev, generator=g),
        (2, 2): torch.randint(0, 1000, (n,), dtype=torch.int64, device=dev, generator=g)}
heavy = int(torch.bincount(z[: min(n, 20_000_000)]).max()) / min(n, 20_000_000)
del z, perm
torch.cuda.synchronize(); print("gen", round(time.time() - t0, 1), "s; heaviest key share", round(heavy, 3), flush=True)
e = qce_b200.Engine(0); lib = bench.host_lib()
for (r, c), t in cols.items():
    e.upload_column_device(r, c, t.data_ptr(), n, adopt=True)
q = "0 1 2|0.1=1.0&1.1=2.0|0.2 1.2 2.2\n"
out = bench.run_query(lib, q)
ok = int(out.split()[0]) == int(cols[(0, 2)].sum())
for _ in range(2): bench.run_query(lib, q)
e.timer_reset()
for _ in range(3): assert bench.run_query(lib, q) == out
ms, launches = e.timer_read()
e.profile(True); bench.run_query(lib, q); prof = {k: v for k, v in e.profile_read().items() if not k.startswith("gap_")}; e.profile(False)
print(json.dumps({"workload": "C4 on ONE GPU: 3 relations x %d rows, R0.c1 ~ Zipf(1.2) over R1's primary keys (heaviest key ~%.0f %% of rows), query %s" % (n, 100 * heavy, q.strip()),
                  "ms_per_query": round(ms / 3, 3), "kernel_launches": int(launches / 3), "input_rows_per_s": 3 * n / (ms / 3 / 1e3), "result": out.strip(),
                  "first_checksum_equals_sum_of_R0_c2": ok,
                  "top_kernels_ms": {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:12]}}))

"""BASELINE config 3 on N GPUs (run under torchrun): 4-way PK-FK join chain with a
filter and a self-join predicate, `rows` rows per relation in total, row-sharded,
every join exchanged over NVLink peer windows (shardexec).  Prints one JSON line.
Full-size check (size-independent properties): the chain is PK-FK, so the result has
exactly one row per filtered+self-joined row of relation 0 and the checksum of 0.3
is the sum of c3 over those rows (computed independently with torch, all-reduced)."""
import os, sys, time, json, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import qce_b200
from qce_b200 import shardexec
from qce_b200.sharded import row_window
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
eng = qce_b200.Engine(lr)
dev = torch.device("cuda", lr)
comm = shardexec.Comm(dist, torch, dev, rank, world)
n = int(float(sys.argv[1])) // (4096 * world) * (4096 * world)   # rows per relation, all ranks together
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
begin, count = row_window(n, rank, world)
gen = torch.Generator(device=dev)
cols, keep = {}, []
a = 2654435761 % n
while math.gcd(a, n) != 1:
    a += 1
for r in range(4):
    gen.manual_seed(300 + 16 * r + rank)
    i = torch.arange(begin, begin + count, dtype=torch.int64, device=dev)
    pk = (i * a + 12345 * (r + 1)) % n                      # a permutation of [0, n): the primary key
    fk = torch.randint(0, n, (count,), dtype=torch.int64, device=dev, generator=gen)
    c2 = torch.where(torch.rand(count, device=dev, generator=gen) < 0.5, fk,
                     torch.randint(0, n, (count,), dtype=torch.int64, device=dev, generator=gen))
    c3 = torch.randint(0, 1000, (count,), dtype=torch.int64, device=dev, generator=gen)
    for c, t in enumerate((pk, fk, c2, c3)):
        cols[(r, c)] = t
    del i
torch.cuda.synchronize()
for (r, c), t in cols.items():
    mx = comm.allreduce_max(eng.column_max_device(t.data_ptr(), count))
    eng.adopt_column_window(r, c, t.data_ptr(), begin, count, n, mx)
shardexec.open_windows(eng, comm, 40 * count + (64 << 20))
ex = shardexec.ShardedExecutor(shardexec.EngineOps(eng), comm)
q = "0 1 2 3|0.1=0.2&0.1=1.0&1.1=2.0&2.1=3.0&0.3<900|0.3 1.3 2.3 3.3"
res = ex.run_query(q)
# independent check of the row count and the first checksum
m = (cols[(0, 3)] < 900) & (cols[(0, 1)] == cols[(0, 2)])
chk = torch.stack([m.sum(), cols[(0, 3)][m].sum()])
dist.all_reduce(chk)
ok = int(chk[0]) == res["pairs"] and int(chk[1]) == res["sums"][0]
for _ in range(2):
    ex.run_query(q)
times = []
for _ in range(steps):
    eng.sync(); torch.cuda.synchronize(); dist.barrier()
    eng.timer_reset()
    r2 = ex.run_query(q)
    ms, launches = eng.timer_read()
    times.append(ms)
    assert r2 == res
t = torch.tensor([sum(times)], dtype=torch.float64, device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms_q = float(t[0]) / steps
eng.profile(True); ex.run_query(q)
prof = {k: round(v["ms"], 3) for k, v in sorted(eng.profile_read().items(), key=lambda kv: -kv[1]["ms"]) if not k.startswith("gap_")}
eng.profile(False)
if rank == 0:
    print(json.dumps({"workload": "C3: 4 relations x %d rows over %d GPUs (row-sharded), query %s" % (n, world, q),
                      "ms_per_query": round(ms_q, 3), "input_rows_per_s": 4 * n / (ms_q / 1e3), "result": shardexec.format_result(res).strip(),
                      "rows_out": res["pairs"], "count_and_first_checksum_match_independent_torch_reduction": ok,
                      "kernel_launches_per_query": int(launches), "bytes_pushed_off_rank_rank0": int(ex.stats["bytes_sent_off_rank"]),
                      "top_kernels_ms_rank0": dict(list(prof.items())[:12])}))
dist.barrier(); dist.destroy_process_group()

"""Phase breakdown of the sharded join (run under torchrun)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import qce_b200
from qce_b200 import sharded
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
eng = qce_b200.Engine(lr)
rows = int(float(sys.argv[1])) // 4096 * 4096
n = world * rows
dev = torch.device("cuda", lr)
gen = torch.Generator(device=dev); cols = {}
for r, seed in enumerate((1, 2)):
    gen.manual_seed(1000 + seed)
    cols[(r, 0)] = torch.arange(n, dtype=torch.int64, device=dev)
    cols[(r, 1)] = torch.randint(0, n, (n,), dtype=torch.int64, device=dev, generator=gen)
    cols[(r, 2)] = torch.randint(0, 10**6, (n,), dtype=torch.int64, device=dev, generator=gen)
torch.cuda.synchronize()
for (r, c), t in cols.items(): eng.upload_column_device(r, c, t.data_ptr(), n, adopt=True)
ops = sharded.EngineOps(eng, torch)
# wrap ops with timers
T = {}
def timed(name, fn):
    def w(*a, **k):
        eng.sync(); torch.cuda.synchronize(); t = time.perf_counter()
        r = fn(*a, **k)
        eng.sync(); torch.cuda.synchronize(); dt = time.perf_counter() - t; T[name] = T.get(name, 0) + dt
        if os.environ.get("QCE_TRACE") and rank == 0:
            rs, us = eng.mempool_stats()
            print(f"   {name:22s} {1e3*dt:8.3f} ms  pool reserved {rs/1e9:7.3f} GB used {us/1e9:7.3f} GB  torch reserved {torch.cuda.memory_reserved()/1e9:.3f}")
        return r
    return w
for name in ["filter_window", "build_from_ids", "build_window", "histogram", "partition", "from_exchange", "sort", "merge_join", "checksum", "release_partition"]:
    setattr(ops, name, timed(name, getattr(ops, name)))
sj = sharded.ShardedJoin(ops, dist, torch, rank, world)
sj._exchange_start = timed("exchange_start(counts + async all_to_all)", sj._exchange_start)
sj._allreduce_u64 = timed("allreduce", sj._allreduce_u64)
spec = sharded.JoinSpec(lhs=(0, 1), rhs=(1, 1), lhs_filter=(2, ">", 500000), lhs_selects=[0], rhs_selects=[0, 2])
for _ in range(3): sj.run(spec, n, n)
T.clear(); eng.profile(True)
K = 5
dist.barrier(); t0 = time.perf_counter()
for _ in range(K): res = sj.run(spec, n, n)
tot = time.perf_counter() - t0
prof = {k: v for k, v in eng.profile_read().items() if not k.startswith("gap_")}
if rank == 0:
    print("total ms/step (with timers)", 1e3 * tot / K, "stats", {k: v for k, v in sj.stats.items() if k != "splitters"})
    for k, v in sorted(T.items(), key=lambda kv: -kv[1]): print(f"  {k:24s} {1e3*v/K:8.3f} ms/step")
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]): print(f"    kernel {k:18s} {v['launches']/K:5.1f} x {v['ms']/K:8.3f} ms/step")
dist.barrier(); dist.destroy_process_group()

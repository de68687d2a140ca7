"""Phase breakdown of the peer-push sharded executor (run under torchrun)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import qce_b200
from qce_b200 import shardexec
from qce_b200.sharded import row_window
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
eng = qce_b200.Engine(lr)
rows = int(float(sys.argv[1])) // 4096 * 4096
n = world * rows
dev = torch.device("cuda", lr)
comm = shardexec.Comm(dist, torch, dev, rank, world)
begin = rank * rows
gen = torch.Generator(device=dev); cols = {}
for r, seed in enumerate((1, 2)):
    gen.manual_seed(1000 + 64 * seed + rank)
    cols[(r, 0)] = torch.arange(begin, begin + rows, dtype=torch.int64, device=dev)
    cols[(r, 1)] = torch.randint(0, n, (rows,), dtype=torch.int64, device=dev, generator=gen)
    cols[(r, 2)] = torch.randint(0, 10**6, (rows,), dtype=torch.int64, device=dev, generator=gen)
torch.cuda.synchronize()
for (r, c), t in cols.items():
    mx = comm.allreduce_max(eng.column_max_device(t.data_ptr(), rows))
    eng.adopt_column_window(r, c, t.data_ptr(), begin, rows, n, mx)
shardexec.open_windows(eng, comm, 24 * rows + (64 << 20))
ops = shardexec.EngineOps(eng)
T = {}
def timed(name, fn):
    def w(*a, **k):
        eng.sync(); t = time.perf_counter()
        r = fn(*a, **k)
        eng.sync(); dt = time.perf_counter() - t; T[name] = T.get(name, 0) + dt
        return r
    return w
for name in ["filter_window", "build_from_ids", "build_window", "histogram", "push_tuples", "push_ids", "ids_hist", "sort", "merge_join", "checksum", "fence"]:
    setattr(ops, name, timed(name, getattr(ops, name)))
for name in ["all_gather_u64", "allreduce_u64", "barrier"]:
    setattr(comm, name, timed("comm." + name, getattr(comm, name)))
ex = shardexec.ShardedExecutor(ops, comm)
q = "0 1|0.1=1.1&0.2>500000|0.0 1.0 1.2"
for _ in range(3): ex.run_query(q)
T.clear(); eng.profile(True)
K = 5
dist.barrier(); t0 = time.perf_counter()
for _ in range(K): res = ex.run_query(q)
tot = time.perf_counter() - t0
prof = {k: v for k, v in eng.profile_read().items() if not k.startswith("gap_")}
if rank == 0:
    print("total ms/step (with timers)", 1e3 * tot / K, "stats", {k: v for k, v in ex.stats.items() if k != "splitters"})
    for k, v in sorted(T.items(), key=lambda kv: -kv[1]): print(f"  {k:24s} {1e3*v/K:8.3f} ms/step")
    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]): print(f"    kernel {k:18s} {v['launches']/K:5.1f} x {v['ms']/K:8.3f} ms/step")
if rank == 1:
    print("rank 1:")
    for k, v in sorted(T.items(), key=lambda kv: -kv[1]): print(f"  r1 {k:24s} {1e3*v/K:8.3f} ms/step")
# un-instrumented steps with a wall-clock trace of when each rank ENTERS / LEAVES its collectives
eng.profile(False)
ops2 = shardexec.EngineOps(eng)
comm3 = shardexec.Comm(dist, torch, dev, rank, world)
trace = []
def traced(name, fn):
    def w(*a, **k):
        t_in = time.perf_counter(); r = fn(*a, **k); trace.append((name, t_in, time.perf_counter())); return r
    return w
for name in ["all_gather_u64", "allreduce_u64", "barrier"]:
    setattr(comm3, name, traced(name, getattr(comm3, name)))
for name in ["filter_window", "build_from_ids", "build_window", "histogram", "push_tuples", "push_ids", "ids_hist", "sort", "merge_join", "checksum", "fence", "tuples_view", "col_view"]:
    setattr(ops2, name, traced("op." + name, getattr(ops2, name)))
ex2 = shardexec.ShardedExecutor(ops2, comm3)
for _ in range(2): ex2.run_query(q)
for step in range(2):
    trace.clear(); eng.sync(); dist.barrier(); t0 = time.perf_counter()
    ex2.run_query(q)
    t_end = time.perf_counter()
    for r in range(world):
        dist.barrier()
        if rank == r:
            print(f"rank {r} step {step}: total {1e3*(t_end-t0):.3f} ms")
            print("   " + " | ".join(f"{n} {1e3*(a-t0):.2f}-{1e3*(b-t0):.2f}" for n, a, b in trace), flush=True)
# latency of the small host-vector collectives themselves (no skew: back to back after a barrier)
comm2 = shardexec.Comm(dist, torch, dev, rank, world)
v = np.arange(512, dtype=np.uint64)
for name, fn in [("all_gather_u64(512)", lambda: comm2.all_gather_u64(v)), ("barrier", comm2.barrier), ("allreduce_u64(5)", lambda: comm2.allreduce_u64(v[:5]))]:
    for _ in range(5): fn()
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(50): fn()
    dt = (time.perf_counter() - t0) / 50
    if rank == 0: print(f"  collective {name:22s} {1e6*dt:8.1f} us")
dist.barrier(); dist.destroy_process_group()

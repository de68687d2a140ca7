"""Config-5-style batch over N GPUs as replicas (run under torchrun, or plainly for one GPU):
every rank uploads the database, whole queries are dealt round-robin, rank 0 prints the
batch's stdout in input order and the throughput.  Synthetic C1/C5-shaped database
(oracle.workload.gen_small_db scaled) -- the generator is test infrastructure, the
execution path is the host layer + libqce_b200.so."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import qce_b200, bench
from qce_b200 import batch_replicas
from oracle import workload as wl
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 200
db = wl.gen_small_db(seed=2018, scale=scale)
queries = wl.gen_queries(db, nq, seed=5, max_joins=2)
eng = qce_b200.Engine(lr); lib = bench.host_lib()
eng.upload_db(db)
run_one = batch_replicas.host_layer_runner(lib)
out = batch_replicas.run_batch_replicated(run_one, queries, dist, rank, world)   # warm-up + result
if world > 1: dist.barrier()
t0 = time.perf_counter()
out2 = batch_replicas.run_batch_replicated(run_one, queries, dist, rank, world)
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
if rank == 0:
    serial = "".join(run_one(q) for q in queries) if world > 1 else out
    print(json.dumps({"workload": "C5-style batch: %d queries over %d relations (%d rows in total), replicas on %d GPU(s)" %
                      (len(queries), len(db), sum(len(r[0]) for r in db), world),
                      "queries_per_s": len(queries) / dt, "ms_per_batch": 1e3 * dt, "same_as_one_gpu_in_order": out == serial and out2 == serial,
                      "lines": out.count("\n")}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()

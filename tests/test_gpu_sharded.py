"""GPU, 2 ranks over NCCL (skipped with fewer than 2 GPUs): the sharded join
against the oracle and against the unsharded engine."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import qce_b200
from qce_b200 import sharded
from oracle import workload as wl, qce_oracle as orc
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
eng = qce_b200.Engine(int(os.environ["LOCAL_RANK"]))
rows = 300_007
db = wl.gen_pair_db(rows, rows // 2, filt_domain=1000)
eng.upload_db(db)
spec = sharded.JoinSpec(lhs=(0, 1), rhs=(1, 1), lhs_filter=(2, ">", 500), lhs_selects=[0], rhs_selects=[0, 2])
sj = sharded.ShardedJoin(sharded.EngineOps(eng, torch), dist, torch, rank, world)
res = sj.run(spec, rows, rows)
res2 = sj.run(spec, rows, rows)
if rank == 0:
    want = orc.run_batch(db, "0 1|0.1=1.1&0.2>500|0.0 1.0 1.2\n")
    print("RESULT " + json.dumps({"ok": sharded.format_result(res) == want and res == res2, "pairs": res["pairs"],
                                  "sent": sj.stats["tuples_sent_off_rank"], "local": sj.stats["local_join_input"]}))
dist.barrier()
dist.destroy_process_group()
''' % ROOT


def test_two_rank_sharded_join_matches_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import tempfile
    script = os.path.join(tempfile.mkdtemp(), "sharded_child.py")
    open(script, "w").write(CHILD)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", script],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert p.returncode == 0 and lines, p.stderr[-3000:]
    r = json.loads(lines[-1][7:])
    assert r["ok"] and r["pairs"] > 0 and r["sent"] > 0


CHILD_EXEC = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
import qce_b200
from qce_b200 import shardexec
from oracle import workload as wl, qce_oracle as orc
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
eng = qce_b200.Engine(lr)
comm = shardexec.Comm(dist, torch, torch.device("cuda", lr), rank, world)
ok, sent = True, 0
first = True
for kind, rows, queries in [("pair", 300_007, ["0 1|0.1=1.1&0.2>500|0.0 1.0 1.2", "0 1|0.1=1.1|0.0 1.2", "0|0.2<100&0.1>3000|0.2"]),
                            ("chain", 400_000, ["0 1 2 3|0.1=0.2&0.1=1.0&1.1=2.0&2.1=3.0&0.3<900|0.3 1.3 2.3 3.3",
                                                "0 1 2|0.1=1.0&0.2=2.0&0.3<200|1.1 2.1 0.0"]),
                            ("zipf", 200_000, ["0 1 2|0.1=1.0&1.1=2.0|0.2 1.2 2.2"])]:
    db = {"pair": lambda: wl.gen_pair_db(rows, rows // 3, filt_domain=1000), "chain": lambda: wl.gen_chain_db(rows),
          "zipf": lambda: wl.gen_zipf_db(rows)}[kind]()
    keep = []
    shardexec.load_sharded_columns(eng, torch, comm, db, keep)
    if first:
        shardexec.open_windows(eng, comm, 1 << 30)
        first = False
    ex = shardexec.ShardedExecutor(shardexec.EngineOps(eng), comm)
    for q in queries:
        got = shardexec.format_result(ex.run_query(q))
        again = shardexec.format_result(ex.run_query(q))
        sent += ex.stats.get("bytes_sent_off_rank", 0)
        want = wl.truth_query(orc.parse_query(q), db)
        if got != want or again != want:
            ok = False
            if rank == 0: print("MISMATCH", q, got, want)
if rank == 0:
    print("RESULT " + json.dumps({"ok": ok, "sent": int(sent)}))
dist.barrier()
dist.destroy_process_group()
''' % ROOT


def test_two_rank_peer_push_executor_matches_truth():
    """2 GPUs: row-sharded columns, tuples / bystander columns / row ids pushed into the
    peer's window over NVLink, chains with carried join keys."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import tempfile
    script = os.path.join(tempfile.mkdtemp(), "exec_child.py")
    open(script, "w").write(CHILD_EXEC)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", script],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    assert p.returncode == 0 and lines, (p.stdout[-2000:], p.stderr[-3000:])
    r = json.loads(lines[-1][7:])
    assert r["ok"] and r["sent"] > 0

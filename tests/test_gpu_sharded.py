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

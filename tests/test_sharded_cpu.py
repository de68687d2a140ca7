"""CPU, world_size 2, gloo: the multi-GPU orchestration (row windows, global
histogram -> splitters, partition bookkeeping, all-to-all, checksum reduce) with a
numpy stand-in for the engine.  The stand-in is test infrastructure (it uses the
oracle's primitives); the product's backend is EngineOps (GPU only)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import qce_oracle as orc  # noqa: E402
from oracle import workload as wl  # noqa: E402


class NumpyOps:
    """Same interface as sharded.EngineOps, on host arrays."""
    comm_device = torch.device("cpu")

    def __init__(self, db):
        self.db = db

    def key_bits(self, rel, col):
        return max(1, int(self.db[rel][col].max()).bit_length())

    def filter_window(self, rel, col, op, c, begin, count):
        return orc.filter_scan(self.db[rel][col][begin:begin + count], op, c) + np.uint64(begin)

    def build_from_ids(self, rel, col, ids):
        k, p = orc.build_tuples(self.db[rel][col], ids)
        return (k << np.uint64(32)) | p

    def build_window(self, rel, col, begin, count):
        k = self.db[rel][col][begin:begin + count]
        return (k << np.uint64(32)) | np.arange(begin, begin + count, dtype=np.uint64)

    def histogram(self, t, key_bits):
        d = ((t >> np.uint64(32)) >> np.uint64(max(key_bits - 8, 0))) & np.uint64(255)
        return np.bincount(d.astype(np.int64), minlength=256).astype(np.uint64)

    def partition(self, t, key_bits, splitters, nparts):
        part = np.searchsorted(np.array(splitters, dtype=np.uint64), t >> np.uint64(32), side="right") \
            if nparts > 1 else np.zeros(len(t), dtype=np.int64)
        order = np.argsort(part, kind="stable")
        counts = [int((part == p).sum()) for p in range(nparts)]
        return counts, torch.from_numpy(t[order].view(np.int64).copy()), None

    def release_partition(self, buf):
        pass

    def from_exchange(self, recv, key_bits, id_bound=0, key_range=(0, 0)):
        return recv.numpy().view(np.uint64).copy()

    def sort(self, t):
        t[:] = t[np.argsort(t >> np.uint64(32), kind="stable")]

    def merge_join(self, L, R):
        m = np.uint64(0xFFFFFFFF)
        return orc.merge_join(L >> np.uint64(32), L & m, R >> np.uint64(32), R & m)

    def checksum(self, ids, rel, cols):
        return [orc.checksum(self.db[rel][c], ids) for c in cols]

    def count(self, ids):
        return len(ids)

    def tuples_count(self, t):
        return len(t)

    def free_ids(self, h):
        pass

    def free_tuples(self, h):
        pass


def _worker(rank, world, port, rows, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import qce_b200  # noqa: F401
    from qce_b200 import sharded
    db = wl.gen_pair_db(rows, rows // 3, filt_domain=1000)
    spec = sharded.JoinSpec(lhs=(0, 1), rhs=(1, 1), lhs_filter=(2, ">", 500), lhs_selects=[0], rhs_selects=[0, 2])
    sj = sharded.ShardedJoin(NumpyOps(db), dist, torch, rank, world)
    res = sj.run(spec, rows, rows)
    if rank == 0:
        out.put((sharded.format_result(res), res["pairs"], sj.stats["splitters"]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_join_matches_oracle(world):
    rows = 20011
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, out)) for r in range(world)]
    for p in procs:
        p.start()
    line, pairs, splitters = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    db = wl.gen_pair_db(rows, rows // 3, filt_domain=1000)
    assert line == orc.run_batch(db, "0 1|0.1=1.1&0.2>500|0.0 1.0 1.2\n")
    assert pairs > 0 and len(splitters) == world - 1 and splitters == sorted(splitters)


def test_choose_splitters_balances():
    import qce_b200  # noqa: F401
    from qce_b200 import sharded
    hist = np.zeros(256, dtype=np.uint64)
    hist[:100] = 10
    sp = sharded.choose_splitters(hist, 27, 4)
    assert sp == [25 << 19, 50 << 19, 75 << 19]
    hist[:] = 0
    hist[7] = 1000  # one heavy bin: cannot be split, the later parts are empty
    sp = sharded.choose_splitters(hist, 16, 4)
    assert sp == sorted(sp) and all(s >= (8 << 8) for s in sp)
    assert sharded.row_window(10_000, 1, 4) == (4096, 4096) and sharded.row_window(10_000, 3, 4) == (10_000, 0)


# ---------------------------------------------------------------- peer-push executor (shardexec)
EXEC_QUERIES = {
    "pair": ["0 1|0.1=1.1&0.2>500|0.0 1.0 1.2", "0 1|0.1=1.1|0.0 1.2", "0|0.2<100|0.0 0.1", "0|0.2<100&0.1>3000|0.2",
             "0 1|0.1=1.1&0.2>999999|0.0 1.0", "1 0|0.1=1.1&1.2<300|0.2 1.0"],
    "chain": ["0 1 2 3|0.1=0.2&0.1=1.0&1.1=2.0&2.1=3.0&0.3<900|0.3 1.3 2.3 3.3",
              "0 1 2|0.1=1.0&1.1=2.0&0.3<500|0.0 1.0 2.3", "0 1 2|0.1=1.0&0.2=2.0|0.3 1.3 2.3",
              "0 1 2 3|0.1=1.0&1.1=2.0&2.1=3.0|0.3 3.3", "0 1 2|0.1=1.0&0.2=2.0&0.3<200|1.1 2.1 0.0"],
    "zipf": ["0 1 2|0.1=1.0&1.1=2.0|0.2 1.2 2.2", "0 1 2|0.1=1.0&1.1=2.0&0.2<500|0.0 2.1"],
}


def _exec_db(kind, rows):
    if kind == "pair":
        return wl.gen_pair_db(rows, rows // 3, filt_domain=1000)
    if kind == "chain":
        return wl.gen_chain_db(rows)
    return wl.gen_zipf_db(rows)


def _exec_worker(rank, world, port, kind, rows, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import qce_b200  # noqa: F401
    from qce_b200 import shardexec
    from tests.numpy_shard_ops import NumpyShardOps
    db = _exec_db(kind, rows)
    comm = shardexec.Comm(dist, torch, torch.device("cpu"), rank, world)
    ex = shardexec.ShardedExecutor(NumpyShardOps(db, rank, world), comm)
    lines, refused = [], 0
    sent = 0
    for q in EXEC_QUERIES[kind]:
        lines.append(shardexec.format_result(ex.run_query(q)))
        sent += ex.stats.get("bytes_sent_off_rank", 0)
    for bad in ["0 1|0.1=1.1&0.2<20&1.2<20|0.0 1.0", "0|0.1=0.2|0.0", "0 0|0.1=1.1|0.0 1.0", "0 1 0|0.1=1.1&1.2=2.1&0.2=2.1|0.0"]:
        try:
            ex.run_query(bad)
        except shardexec.UnsupportedQuery:
            refused += 1
            assert ex.stats == {}  # refused while planning: nothing ran, nothing is left on the device
    if rank == 0:
        out.put((lines, refused, sent))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,rows", [(2, "pair", 20011), (3, "pair", 9001), (2, "chain", 12000), (3, "chain", 20000),
                                             (2, "zipf", 15000), (1, "chain", 5000),
                                             (3, "chain", 3000),    # fewer rows than one 4096-row window: ranks 1, 2 own nothing
                                             (3, "zipf", 9000)])
def test_sharded_executor_matches_truth(world, kind, rows):
    """Row-sharded columns, pushes into peer windows, bystander columns by slot, carried join
    keys, projection by id push: checksums equal the relational truth (= the reference inside PDQ-T)."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 31500 + os.getpid() % 2000 + 7 * world + len(kind)
    procs = [ctx.Process(target=_exec_worker, args=(r, world, port, kind, rows, out)) for r in range(world)]
    for p in procs:
        p.start()
    lines, refused, sent = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    db = _exec_db(kind, rows)
    for q, line in zip(EXEC_QUERIES[kind], lines):
        assert line == wl.truth_query(orc.parse_query(q), db), q
    assert refused == 4
    assert (sent > 0) == (world > 1)
    if kind == "pair":  # also the reference's own answer (oracle restatement), byte for byte
        assert lines[0] == orc.run_batch(db, EXEC_QUERIES[kind][0] + "\n")


# ---------------------------------------------------------------- query-level replicas (config 5)
def _replica_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import qce_b200  # noqa: F401
    from qce_b200 import batch_replicas
    from tests.helpers import load_db, load_json
    db = load_db("ops_db.npz")
    text = "".join(r["query"] + "\nF\n" for r in load_json("ops.json") if r["class"] == "PDQ-T")
    queries = batch_replicas.split_queries(text)
    calls = []

    def run_one(q):
        calls.append(q)
        return orc.run_batch(db, q + "\n")
    res = batch_replicas.run_batch_replicated(run_one, queries, dist, rank, world)
    if rank == 0:
        out.put((res, len(queries), len(calls)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 3])
def test_replicated_batch_keeps_input_order(world):
    """Whole queries dealt round-robin over ranks; rank 0 emits every query's block (count
    lines + result line) in input order = the reference's recorded stdout."""
    from tests.helpers import load_json
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 33500 + os.getpid() % 2000 + world
    procs = [ctx.Process(target=_replica_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res, nq, ncalls = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    recs = [r for r in load_json("ops.json") if r["class"] == "PDQ-T"]
    assert res == "".join(r["stdout"] for r in recs)
    assert nq == len(recs) and ncalls == -(-nq // world)

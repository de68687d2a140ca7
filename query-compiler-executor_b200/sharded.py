"""Intra-query sharding of one filter + equi-join + projection over N GPUs
(BASELINE.json north_star: "Large joins shard across the 8 GPUs by
radix high bits, with an NCCL all-to-all over NVLink and a final checksum
reduce").  One process per GPU, `torch.distributed` for the collectives; every
loop over tuples is a C-ABI call into libqce_b200.so.

Per join, on every rank:
  1. scan / build tuples for the rank's ROW WINDOW of each input (row ids stay
     relation-global)                      qce_filter_scan_range, qce_build_tuples_*
  2. 256-bin histogram of the top key bits qce_key_histogram  -> all_reduce(SUM)
     -> nparts-1 splitters on bin boundaries that balance tuples per rank
  3. group the run by destination rank     qce_partition_tuples (one-sweep kernel,
                                           range digit; stable)
  4. all_to_all_single: counts, then the packed 8-byte words over NVLink
  5. local sort + merge join of the received key range, local checksums
  6. all_reduce(SUM) of the uint64 checksums and pair counts (exact: addition
     mod 2^64 is associative and commutative)

Base columns are REPLICATED on every GPU in this round (each rank gathers
projected values for the pairs it produced from its own HBM); row-range-sharded
columns with a second exchange are the next step for relations that do not fit.
A single key's group cannot be split by key partitioning: a key heavier than
1/N of the input leaves its rank overloaded.

The engine is reached through a small `ops` object so that the orchestration
(splitters, counts, exchange bookkeeping, reductions) also runs on CPU under the
gloo backend with a numpy stand-in -- tests/test_sharded_cpu.py.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class JoinSpec:
    """`lhs.col = rhs.col`, optional filter on the lhs binding, projected columns."""
    lhs: Tuple[int, int]
    rhs: Tuple[int, int]
    lhs_filter: Optional[Tuple[int, str, int]]  # (column, op, constant) on the lhs relation
    lhs_selects: Sequence[int]
    rhs_selects: Sequence[int]


def choose_splitters(hist: np.ndarray, key_bits: int, nparts: int) -> List[int]:
    """Splitter keys on histogram-bin boundaries so that every part gets about
    total/nparts tuples.  hist = global 256-bin histogram of the top 8 significant
    key bits.  part(key) = #splitters <= key."""
    shift = max(key_bits - 8, 0)
    total = int(hist.sum())
    cum = np.cumsum(hist.astype(np.int64))
    out = []
    for k in range(1, nparts):
        target = total * k / nparts
        b = int(np.searchsorted(cum, target, side="left")) + 1  # first bin of the next part
        b = min(max(b, (out[-1] >> shift) if out else 0), 256)
        out.append(b << shift)
    return out


def row_window(rows: int, rank: int, world: int, align: int = 4096) -> Tuple[int, int]:
    """Rank's row range; boundaries aligned so vector loads stay aligned."""
    per = -(-rows // world)
    per = -(-per // align) * align
    begin = min(rank * per, rows)
    return begin, min(per, rows - begin)


class EngineOps:
    """GPU backend: qce_b200.Engine + torch CUDA tensors for the exchange buffers."""

    def __init__(self, engine, torch):
        self.e, self.torch = engine, torch
        self.comm_device = torch.device("cuda", torch.cuda.current_device())

    class _DevArray:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}

    def key_bits(self, rel, col):
        return max(1, int(self.e.column_info(rel, col)[1]).bit_length())

    def filter_window(self, rel, col, op, c, begin, count):
        return self.e.filter_scan(rel, col, op, c, rows=(begin, count))

    def build_from_ids(self, rel, col, ids):
        return self.e.build_tuples(rel, col, ids)

    def build_window(self, rel, col, begin, count):
        return self.e.build_tuples(rel, col, rows=(begin, count))

    def histogram(self, t, key_bits):
        return self.e.key_histogram(t, key_bits)

    def partition(self, t, key_bits, splitters, nparts):
        counts, buf = self.e.partition_tuples(t, key_bits, splitters, nparts)  # returns synchronised
        n = sum(counts)
        send = self.torch.as_tensor(self._DevArray(buf, n), device=self.comm_device) if n else \
            self.torch.empty(0, dtype=self.torch.int64, device=self.comm_device)
        return counts, send, buf

    def release_partition(self, buf):
        self.e.exchange_release(buf)

    def from_exchange(self, recv, key_bits, id_bound=0, key_range=(0, 0)):
        # the run is sorted and joined in place in the receive buffer (no copy); the caller
        # keeps `recv` alive until the run is freed and has waited for the all-to-all
        self.torch.cuda.current_stream().synchronize()
        return self.e.tuples_from_device_packed(recv.data_ptr() if recv.numel() else 0, recv.numel(), key_bits,
                                                id_bound, key_range[0], key_range[1], adopt=recv.numel() > 0)

    def sort(self, t):
        self.e.sort_tuples(t)

    def merge_join(self, L, R):
        return self.e.merge_join(L, R)

    def checksum(self, ids, rel, cols):
        return self.e.checksum(ids, rel, list(cols)) if cols else []

    def count(self, ids):
        return self.e.rowids_count(ids)

    def free_ids(self, h):
        self.e.rowids_free(h)

    def free_tuples(self, h):
        self.e.tuples_free(h)

    def tuples_count(self, t):
        return self.e.tuples_count(t)


class ShardedJoin:
    def __init__(self, ops, dist, torch, rank: int, world: int):
        self.ops, self.dist, self.torch, self.rank, self.world = ops, dist, torch, rank, world
        self.dev = ops.comm_device
        self.stats = {}

    # ---- collectives on small host vectors (wrap modulo 2^64 through int64)
    def _allreduce_u64(self, values: np.ndarray) -> np.ndarray:
        t = self.torch.from_numpy(values.astype(np.uint64).view(np.int64).copy()).to(self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy().view(np.uint64)

    def _exchange_start(self, counts: List[int], send):
        """Counts first (small, blocking), then the payload all-to-all as an async
        op so that the next run's partition kernel overlaps the transfer."""
        torch, dist = self.torch, self.dist
        sc = torch.tensor(counts, dtype=torch.int64, device=self.dev)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc)
        rcounts = [int(x) for x in rc.cpu().tolist()]
        recv = torch.empty(sum(rcounts), dtype=torch.int64, device=self.dev)
        work = dist.all_to_all_single(recv, send, output_split_sizes=rcounts, input_split_sizes=counts, async_op=True)
        return recv, work

    def run(self, spec: JoinSpec, rows_lhs: int, rows_rhs: int) -> dict:
        ops, world, rank = self.ops, self.world, self.rank
        t0 = time.perf_counter()
        lb, lc = row_window(rows_lhs, rank, world)
        rb, rc = row_window(rows_rhs, rank, world)
        # 1. this rank's share of scan + tuple build
        if spec.lhs_filter is not None:
            fcol, fop, fconst = spec.lhs_filter
            ids = ops.filter_window(spec.lhs[0], fcol, fop, fconst, lb, lc)
            L = ops.build_from_ids(spec.lhs[0], spec.lhs[1], ids)
            ops.free_ids(ids)
        else:
            L = ops.build_window(spec.lhs[0], spec.lhs[1], lb, lc)
        R = ops.build_window(spec.rhs[0], spec.rhs[1], rb, rc)
        key_bits = max(ops.key_bits(*spec.lhs), ops.key_bits(*spec.rhs))
        if key_bits > 32:
            raise NotImplementedError("the sharded exchange carries packed (key < 2^32) runs only")
        # 2. splitters from the global key histogram
        hist = self._allreduce_u64(ops.histogram(L, key_bits) + ops.histogram(R, key_bits))
        splitters = choose_splitters(hist, key_bits, world)
        # this rank's key interval (sizes the buckets of its local sort)
        lo = splitters[rank - 1] if rank > 0 else 0
        hi = (splitters[rank] - 1) if rank < world - 1 else (1 << key_bits) - 1
        my_range = (lo, max(lo, hi))
        # 3+4. group by destination, exchange
        t1 = time.perf_counter()
        sent = 0
        pending = []
        for run, nrows in ((L, rows_lhs), (R, rows_rhs)):
            counts, send, buf = ops.partition(run, key_bits, splitters, world)
            sent += sum(counts) - counts[rank]
            recv, work = self._exchange_start(counts, send)
            pending.append((recv, work, send, buf, nrows))
            ops.free_tuples(run)
        recv_runs, keep_alive = [], []
        for recv, work, send, buf, nrows in pending:
            work.wait()
            recv_runs.append(ops.from_exchange(recv, key_bits, nrows, my_range))
            ops.release_partition(buf)
            keep_alive.append(recv)  # the runs live in these buffers until they are freed
            del send
        t2 = time.perf_counter()
        # 5. local sort + merge + checksums of this rank's key range
        L2, R2 = recv_runs
        ops.sort(L2)
        ops.sort(R2)
        oL, oR = ops.merge_join(L2, R2)
        pairs = ops.count(oL)
        sums = ops.checksum(oL, spec.lhs[0], spec.lhs_selects) + ops.checksum(oR, spec.rhs[0], spec.rhs_selects)
        local_in = ops.tuples_count(L2) + ops.tuples_count(R2)
        for h in (oL, oR):
            ops.free_ids(h)
        ops.free_tuples(L2)
        ops.free_tuples(R2)
        del keep_alive
        # 6. reduce
        red = self._allreduce_u64(np.array(sums + [pairs, local_in], dtype=np.uint64))
        t3 = time.perf_counter()
        self.stats = {"build_s": t1 - t0, "exchange_s": t2 - t1, "local_s": t3 - t2, "tuples_sent_off_rank": int(sent),
                      "local_join_input": int(local_in), "splitters": splitters}
        return {"sums": [int(x) for x in red[:len(sums)]], "pairs": int(red[len(sums)]),
                "join_input_tuples": int(red[len(sums) + 1])}


def format_result(res: dict) -> str:
    """The line the reference prints for the same query (print_sums,
    /root/reference/src/utilities.c:212-223)."""
    return "".join("NULL " if res["pairs"] == 0 else f"{s} " for s in res["sums"]) + "\n"


# ---------------------------------------------------------------------------- bench (N > 1)
def bench(eng, host_lib, dist, rank, world, rows, steps, warmup, verify=True):
    """Weak scaling of config 2: every rank owns a `rows`-row window of two
    relations of world*rows rows (columns replicated, generated on the device with
    the same seed on every rank)."""
    import ctypes as C
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    n = world * rows
    gen = torch.Generator(device=dev)
    cols = {}
    for r, seed in enumerate((1, 2)):
        gen.manual_seed(1000 + seed)
        cols[(r, 0)] = torch.arange(n, dtype=torch.int64, device=dev)
        cols[(r, 1)] = torch.randint(0, n, (n,), dtype=torch.int64, device=dev, generator=gen)
        cols[(r, 2)] = torch.randint(0, 10 ** 6, (n,), dtype=torch.int64, device=dev, generator=gen)
    torch.cuda.synchronize()
    for (r, c), t in cols.items():
        eng.upload_column_device(r, c, t.data_ptr(), n, adopt=True)
    spec = JoinSpec(lhs=(0, 1), rhs=(1, 1), lhs_filter=(2, ">", 500000), lhs_selects=[0], rhs_selects=[0, 2])
    sj = ShardedJoin(EngineOps(eng, torch), dist, torch, rank, world)

    def sync_all():
        eng.sync()
        torch.cuda.synchronize()
        dist.barrier()

    res = sj.run(spec, n, n)
    want = format_result(res)
    verified = None
    if verify and rank == 0 and n < (1 << 30):
        # the unsharded host layer on the same (replicated) columns, once, untimed
        buf = C.create_string_buffer(4096)
        failed = C.c_int(0)
        host_lib.qce_host_run_batch(b"0 1|0.1=1.1&0.2>500000|0.0 1.0 1.2\n", buf, 4096, C.byref(failed))
        verified = (buf.value.decode() == want) and failed.value == 0
    for _ in range(warmup):
        sj.run(spec, n, n)
    sync_all()
    times, ex_times = [], []
    for _ in range(steps):
        sync_all()
        eng.timer_reset()
        res = sj.run(spec, n, n)
        ms, _ = eng.timer_read()
        torch.cuda.synchronize()
        times.append(ms)
        ex_times.append(sj.stats["exchange_s"] * 1e3)
        assert format_result(res) == want
    launches = eng.timer_read()[1]
    t = torch.tensor([sum(times), sum(ex_times)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, ex_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / steps
    sent = torch.tensor([sj.stats["tuples_sent_off_rank"]], dtype=torch.float64, device=dev)
    dist.all_reduce(sent, op=dist.ReduceOp.MAX)
    nvlink_gbs = float(sent[0]) * 8 / (ex_ms / steps / 1e3) / 1e9 if ex_ms else None

    # ---- e2e: every step each rank first copies ITS row window of the six
    # referenced columns from pinned host memory and the replicas are rebuilt with
    # an all-gather over NVLink; the checksums come back to the host
    lb, lc = row_window(n, rank, world)
    pinned = {k: v[lb:lb + lc].cpu().pin_memory() for k, v in cols.items()}
    e2e_times = []
    for i in range(min(warmup, 1) + steps):
        sync_all()
        eng.timer_reset()
        for k, full in cols.items():
            full[lb:lb + lc].copy_(pinned[k], non_blocking=True)
        for k, full in cols.items():
            dist.all_gather_into_tensor(full, full[lb:lb + lc])
        torch.cuda.synchronize()
        res = sj.run(spec, n, n)
        ms, _ = eng.timer_read()
        if i >= min(warmup, 1):
            e2e_times.append(ms)
        assert format_result(res) == want
    te = torch.tensor([sum(e2e_times)], dtype=torch.float64, device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0]) / steps
    return {
        "value": 2.0 * n / (ms_per_step / 1e3), "ms_per_step": ms_per_step,
        "config": {"workload": "C2 weak-scaled: 2-way equi-join + range filter, 2 x %d-row uint64 relations "
                               "(%d rows per GPU per relation), query 0 1|0.1=1.1&0.2>500000|0.0 1.0 1.2, sharded by "
                               "key range over %d GPUs (histogram all-reduce, NCCL all-to-all, checksum all-reduce); "
                               "base columns replicated" % (n, rows, world),
                   "rows_per_relation": n, "result": want.strip(), "verified_against_unsharded_engine": verified,
                   "l2": "inputs larger than L2"},
        "exchange": {"ms_per_step_incl_partition": ex_ms / steps, "max_bytes_sent_per_rank": float(sent[0]) * 8,
                     "effective_gbs_per_rank_incl_partition": nvlink_gbs, "nvlink_peak_gbs": 770.0},
        "gpu_launches": int(launches) * steps,
        "e2e": {"value": 2.0 * n / (e2e_ms / 1e3), "unit": "rows/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 6 * lc * 8, "d2h_bytes_per_step": 5 * 8,
                "note": "per rank: own row window from pinned host memory, replicas rebuilt by all-gather over NVLink"},
    }

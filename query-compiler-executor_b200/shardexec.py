"""Sharded execution of one query over N GPUs with peer-memory exchanges
(BASELINE.json north_star: "Large joins shard across the 8 GPUs by
radix high bits ... over NVLink and a final checksum reduce").

One process per GPU.  Base columns are ROW-SHARDED (each rank holds one row
window of every relation; nothing is replicated), intermediate results are
sharded by the key range of the last join.  Per join:

  1. every rank builds the tuple runs of its share: the entity side from its
     slice of the intermediate result (or a filter output / base window), the new
     relation from its base row window                     (C-ABI build calls)
  2. 256-bin histograms of the top key bits, ONE all-gather of both -> every rank
     derives the same splitters (bytes balanced per rank) and the whole
     count matrix [src][dst], hence every segment offset in every window
  3. qce_push_tuples: the partition pass stores each tuple directly into the
     owner's receive window over NVLink (P2P stores; no send buffer, no
     all-to-all).  Row-id columns of the other bindings of the entity and join
     keys needed later follow their tuple with qce_push_u32_by_slot
  4. barrier; local sort + merge join of the received key range; bystander
     columns re-aligned by position (qce_rowids_gather)
Projection (print_sums, src/utilities.c:197-224): the row ids of each projected
binding are pushed to the rank that owns the rows (grouped by row region on
arrival), gathered and summed there, and the uint64 sums are all-reduced (exact:
addition mod 2^64 is associative and commutative).

Query class: left-deep join trees -- filters on at most one binding, self-join
predicates on that binding, then joins that each add one fresh base relation.
Inside this class the reference is tie-invariant and equals relational truth
(PDQ-T); everything else is refused with UnsupportedQuery (never
answered differently from the reference).  Keys must be < 2^32 (packed runs).

The engine is reached through an `ops` object so that the orchestration runs on
CPU under gloo with a numpy stand-in (tests/test_sharded_cpu.py).
"""
from __future__ import annotations

import re
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .engine import exchange_plan, load_library, rowid_push_plan
from .sharded import row_window

_LIB = None


def _lib():
    """libqce_b200.so: the exchange bookkeeping is host arithmetic inside the C-ABI library
    (qce_exchange_plan / qce_rowid_push_plan); it needs the library, not a device."""
    global _LIB
    if _LIB is None:
        _LIB = load_library()
    return _LIB


class UnsupportedQuery(NotImplementedError):
    pass


# ------------------------------------------------------------------ query text
def parse_query(text: str):
    """`r r r|p&p&p|s s` (src/parsing.c:4-148): relations, predicates, selects."""
    rels_s, preds_s, sels_s = text.strip().split("|")
    relations = [int(x) for x in rels_s.split()]
    filters, joins = [], []
    for p in preds_s.split("&"):
        m = re.fullmatch(r"\s*(\d+)\.(\d+)\s*([=<>])\s*(\d+)(?:\.(\d+))?\s*", p)
        if not m:
            raise ValueError(f"bad predicate {p!r}")
        b, c, op, x, y = int(m[1]), int(m[2]), m[3], int(m[4]), m[5]
        if y is None:
            filters.append((b, c, op, x))
        else:
            if op != "=":
                raise ValueError("joins are equi-joins")
            joins.append(((b, c), (x, int(y))))
    selects = []
    for s in sels_s.split():
        b, c = s.split(".")
        selects.append((int(b), int(c)))
    return relations, filters, joins, selects


# ------------------------------------------------------------------ collectives
class Comm:
    """Small host-vector collectives on torch.distributed (NCCL on GPUs, gloo on CPU)."""

    def __init__(self, dist, torch, device, rank: int, world: int):
        self.dist, self.torch, self.dev, self.rank, self.world = dist, torch, device, rank, world
        self.n_collectives = 0

    def all_gather_u64(self, v: np.ndarray) -> np.ndarray:
        t = self.torch.from_numpy(np.ascontiguousarray(v, dtype=np.uint64).view(np.int64).copy()).to(self.dev)
        out = self.torch.empty(self.world * t.numel(), dtype=self.torch.int64, device=self.dev)
        self.dist.all_gather_into_tensor(out, t)
        self.n_collectives += 1
        return out.cpu().numpy().view(np.uint64).reshape(self.world, -1)

    def allreduce_u64(self, v: np.ndarray) -> np.ndarray:
        t = self.torch.from_numpy(np.ascontiguousarray(v, dtype=np.uint64).view(np.int64).copy()).to(self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        self.n_collectives += 1
        return t.cpu().numpy().view(np.uint64)

    def allreduce_max(self, v: int) -> int:
        t = self.torch.tensor([v], dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return int(t.cpu()[0])

    def barrier(self) -> None:
        t = self.torch.zeros(1, dtype=self.torch.int64, device=self.dev)
        self.dist.all_reduce(t)
        self.n_collectives += 1
        t.cpu()

    def all_gather_bytes(self, b: bytes) -> List[bytes]:
        t = self.torch.frombuffer(bytearray(b), dtype=self.torch.uint8).to(self.dev)
        out = self.torch.empty(self.world * len(b), dtype=self.torch.uint8, device=self.dev)
        self.dist.all_gather_into_tensor(out, t)
        raw = bytes(out.cpu().numpy().tobytes())
        return [raw[i * len(b):(i + 1) * len(b)] for i in range(self.world)]


# ------------------------------------------------------------------ GPU backend
class EngineOps:
    """qce_b200.Engine behind the executor's op interface: one C-ABI call each."""

    def __init__(self, engine):
        self.e = engine

    def key_bits(self, rel, col):
        return max(1, int(self.e.column_info(rel, col)[1]).bit_length())

    def key_max(self, rel, col):
        return int(self.e.column_info(rel, col)[1])

    def rows(self, rel, col=0):
        return int(self.e.column_info(rel, col)[0])

    def filter_window(self, rel, col, op, c, begin, count):
        return self.e.filter_scan(rel, col, op, c, rows=(begin, count))

    def filter_refine(self, ids, rel, col, op, c):
        self.e.filter_refine(ids, rel, col, op, c)
        return ids

    def self_join(self, rel, c1, c2, ids):
        a, b = self.e.scan_join(rel, c1, ids, rel, c2, ids)
        self.e.rowids_free(b)
        return a

    def build_from_ids(self, rel, col, ids):
        return self.e.build_tuples(rel, col, ids)

    def build_window(self, rel, col, begin, count):
        return self.e.build_tuples(rel, col, rows=(begin, count))

    def narrow_window(self, rel, col, begin, count):
        return self.e.column_window_u32(rel, col, begin, count)

    def gather_column(self, rel, col, ids):
        return self.e.column_gather_u32(rel, col, ids)

    def iota(self, begin, count, bound):
        return self.e.rowids_iota(begin, count, bound)

    def tuples_from_u32(self, keys, key_bits):
        return self.e.tuples_from_u32(keys, key_bits)

    def histogram(self, t, key_bits):
        return self.e.key_histogram(t, key_bits)

    def push_tuples(self, t, key_bits, splitters, nparts, dst_word_offset, dst_run_index, want_slots):
        return self.e.push_tuples(t, key_bits, splitters, nparts, dst_word_offset, dst_run_index, want_slots)

    def push_col(self, col, slots, nparts, dst_u32_offset):
        self.e.push_u32_by_slot(col, slots, nparts, dst_u32_offset)

    def push_tuples_cols(self, t, key_bits, splitters, nparts, dst_word_offset, dst_run_index, cols, col_u32_offset):
        self.e.push_tuples_cols(t, key_bits, splitters, nparts, dst_word_offset, dst_run_index, cols, col_u32_offset)

    def ids_hist(self, ids, per, width, bpr, world):
        return self.e.rowids_bin_histogram(ids, per, width, bpr, world)

    def push_ids(self, ids, per, width, bpr, world, bin_u32_offset):
        self.e.push_rowids(ids, per, width, bpr, world, bin_u32_offset)

    def fence(self):
        self.e.sync()

    def tuples_view(self, word_offset, n, key_bits, id_bound, key_range):
        return self.e.tuples_from_window(word_offset, n, key_bits, id_bound, key_range[0], key_range[1])

    def col_view(self, u32_offset, n, id_bound=0, bucketed=False, id_min=0):
        return self.e.rowids_from_window(u32_offset, n, id_bound, bucketed, id_min)

    def sort(self, t):
        self.e.sort_tuples(t)

    def merge_join(self, L, R):
        return self.e.merge_join(L, R)

    def gather(self, col, index):
        return self.e.rowids_gather(col, index)

    def checksum(self, ids, rel, cols):
        return self.e.checksum(ids, rel, list(cols)) if cols else []

    def count(self, ids):
        return self.e.rowids_count(ids)

    def tuples_count(self, t):
        return self.e.tuples_count(t)

    def free_ids(self, h):
        self.e.rowids_free(h)

    def free_tuples(self, h):
        self.e.tuples_free(h)

    def window_bytes(self):
        return self.e.xwin_info()[0]


class PeerWindowsUnavailable(RuntimeError):
    """Some rank could not create or map the receive windows (no P2P / CUDA IPC between
    the GPUs of this box): raised on EVERY rank, so callers can switch transport together."""


def open_windows(engine, comm: Comm, nbytes: int) -> None:
    """Create this rank's receive window and map every peer's (CUDA IPC handles
    all-gathered once).  world == 1: the window is only ever addressed locally."""
    err = ""
    try:
        handle = engine.xwin_create(nbytes)
    except Exception as e:  # noqa: BLE001 -- the outcome is agreed on collectively below
        handle, err = bytes(64), repr(e)
    if comm.world > 1:
        handles = comm.all_gather_bytes(handle)
        if not err:
            try:
                engine.xwin_attach(comm.world, comm.rank, b"".join(handles))
            except Exception as e:  # noqa: BLE001
                err = repr(e)
    if comm.allreduce_max(1 if err else 0):
        engine.xwin_destroy()
        raise PeerWindowsUnavailable(err or "a peer rank could not map the windows")


def load_sharded_columns(engine, torch, comm: Comm, db: Sequence[Sequence[np.ndarray]], keep: list) -> None:
    """Each rank uploads only its row window of every column (global row ids);
    the column maximum (sort statistics) is all-reduced."""
    dev = torch.device("cuda", torch.cuda.current_device())
    for r, cols in enumerate(db):
        n = len(cols[0])
        begin, count = row_window(n, comm.rank, comm.world)
        for c, v in enumerate(cols):
            t = torch.from_numpy(np.ascontiguousarray(v[begin:begin + count], dtype=np.uint64).view(np.int64).copy()).to(dev)
            keep.append(t)
            mx = comm.allreduce_max(int(v[begin:begin + count].max()) if count else 0)
            engine.adopt_column_window(r, c, t.data_ptr() if count else 0, begin, count, n, mx)


# ------------------------------------------------------------------ executor
def _align(x: int, a: int = 16) -> int:
    return -(-x // a) * a


class ShardedExecutor:
    def __init__(self, ops, comm: Comm, replicated: bool = False):
        self.ops, self.comm, self.rank, self.world = ops, comm, comm.rank, comm.world
        self.replicated = replicated  # base columns replicated on every rank: projections gather locally
        self.stats: Dict[str, float] = {}

    # ---- one exchange of the two inputs of a join -----------------------------
    def _exchange_pair(self, sides, key_bits: int, key_max: int):
        """sides = [(run, [u32 columns...]), (run, [...])].  Returns, per side, the
        received run (a view of this rank's window) and the received columns."""
        ops, world, me = self.ops, self.world, self.rank
        t0 = time.perf_counter()
        hists = np.concatenate([ops.histogram(run, key_bits) for run, _ in sides])
        H = self.comm.all_gather_u64(hists).reshape(world, len(sides), 256)
        # splitters, count matrix and window layout: the same arithmetic on every rank (C-ABI)
        ncols = [len(cols) for _, cols in sides]
        splitters, recv, before, run_off, col_flat, need, sent_tuples = exchange_plan(_lib(), H, world, me, ncols, key_bits)
        col_off, at = [], 0
        for n in ncols:
            col_off.append([col_flat[at + j] for j in range(n)])
            at += n
        cap = ops.window_bytes()
        if need > cap:
            raise MemoryError(f"exchange needs {need} bytes of receive window, {cap} available")
        sent = 0
        for k, (run, cols) in enumerate(sides):
            dst_words = run_off[k] // np.uint64(8) + before[k]
            if not cols:    # payloads are row ids and travel inside the tuples
                ops.push_tuples(run, key_bits, splitters, world, dst_words, None, False)
            else:           # payload := index in the receiver's run; the columns follow in the same kernel
                for j0 in range(0, len(cols), 6):
                    regions = np.stack([col_off[k][j] // np.uint64(4) + before[k] for j in range(j0, min(j0 + 6, len(cols)))])
                    ops.push_tuples_cols(run, key_bits, splitters, world, dst_words, before[k].astype(np.uint32),
                                         cols[j0:j0 + 6], regions.astype(np.uint64))
            sent += int(sent_tuples[k]) * (8 + 4 * len(cols))
        ops.fence()
        self.comm.barrier()  # every peer's stores into this rank's window have completed
        lo = splitters[me - 1] if me > 0 else 0
        # the last rank's range ends at the largest key that exists, not at 2^key_bits - 1: the
        # local sort sizes its MSD buckets from this range (sparse buckets overflow the finish)
        hi = (splitters[me] - 1) if me < world - 1 else key_max
        out = []
        for k, (_, cols) in enumerate(sides):
            n = int(recv[k][me])
            run = ops.tuples_view(int(run_off[k][me]) // 8, n, key_bits, 0, (lo, max(lo, hi)))
            views = [ops.col_view(int(col_off[k][j][me]) // 4, n) for j in range(len(cols))]
            out.append((run, views))
        self.stats["exchange_s"] = self.stats.get("exchange_s", 0.0) + time.perf_counter() - t0
        self.stats["bytes_sent_off_rank"] = self.stats.get("bytes_sent_off_rank", 0) + sent
        self.stats["splitters"] = splitters
        return out

    # ---- projection -------------------------------------------------------------
    def _project(self, relations, ent: Dict[int, int], selects, nrows: int):
        ops, world, me = self.ops, self.world, self.rank
        by_binding: Dict[int, List[int]] = {}
        for b, c in selects:
            by_binding.setdefault(b, [])
            if c not in by_binding[b]:
                by_binding[b].append(c)
        for b in by_binding:
            if b not in ent:
                raise UnsupportedQuery(f"projected binding {b} takes part in no predicate")
        sums: Dict[Tuple[int, int], int] = {}
        if self.replicated or nrows < 0:
            for b, cols in by_binding.items():
                for c, s in zip(cols, ops.checksum(ent[b], relations[b], cols)):
                    sums[(b, c)] = s
            return sums
        t0 = time.perf_counter()
        geo, hists = {}, []
        # One bin per owner rank: a tile's ids leave in a few long runs (NVLink wants >= 128-byte
        # writes; a 256-bin scatter over the link measured 2.6x slower per id than the same
        # scatter in local HBM).  The owner buckets what it received by L2-sized row regions
        # itself (qce_checksum, window-relative digits) when its window is large enough to pay.
        bpr = 1
        for b in by_binding:
            rows = ops.rows(relations[b])
            per = max(row_window(rows, 0, world)[1], 1)
            width = max(1, -(-per // bpr))
            geo[b] = (per, width, rows)
            hists.append(ops.ids_hist(ent[b], per, width, bpr, world))
        nb = bpr * world
        H = self.comm.all_gather_u64(np.concatenate(hists)).reshape(world, len(by_binding), nb)
        offs, view_off, view_cnt, need, sent_ids = rowid_push_plan(_lib(), H, world, me, len(by_binding), bpr)
        if need > ops.window_bytes():
            raise MemoryError("projection needs more receive window than is available")
        views = {}
        sent = 0
        for k, b in enumerate(by_binding):
            per, width, rows = geo[b]
            ops.push_ids(ent[b], per, width, bpr, world, offs[k])
            views[b] = (int(view_off[k]), int(view_cnt[k]), rows)
            sent += 4 * int(sent_ids[k])
        ops.fence()
        self.comm.barrier()
        for b, cols in by_binding.items():
            off, n, rows = views[b]
            wb, wc = row_window(rows, me, world)
            v = ops.col_view(off, n, wb + wc, bpr > 1, wb)
            for c, s in zip(cols, ops.checksum(v, relations[b], cols)):
                sums[(b, c)] = s
            ops.free_ids(v)
        self.stats["project_exchange_s"] = self.stats.get("project_exchange_s", 0.0) + time.perf_counter() - t0
        self.stats["bytes_sent_off_rank"] = self.stats.get("bytes_sent_off_rank", 0) + sent
        return sums

    # ---- the query ----------------------------------------------------------------
    def run_query(self, text: str) -> dict:
        ops, world, me = self.ops, self.world, self.rank
        self.stats = {}
        relations, filters, joins, selects = parse_query(text)
        fbind = sorted({f[0] for f in filters})
        if len(fbind) > 1:
            raise UnsupportedQuery("filters on more than one binding (reference: positional scan join, SURVEY 8c-ii)")
        self_joins = [j for j in joins if j[0][0] == j[1][0]]
        chain = [j for j in joins if j[0][0] != j[1][0]]
        for (b1, c1), (b2, c2) in chain:
            if relations[b1] == relations[b2] and c1 == c2:
                raise UnsupportedQuery("same relation and column on both sides: the reference skips this join (8c-v)")
        # the whole plan is checked before anything runs: a refusal never leaves device state behind
        for (b, _), _ in self_joins:
            if b not in fbind:
                raise UnsupportedQuery("self-join predicate on a binding without a prior filter (reference exits, 8c-iii)")
        joined, todo = set(fbind), list(chain)
        while todo:
            nxt = next((j for j in todo if not joined or (j[0][0] in joined) != (j[1][0] in joined)), None)
            if nxt is None:
                raise UnsupportedQuery("join between two bindings that are both (or neither) in the intermediate result")
            todo.remove(nxt)
            for b, c in nxt:
                if ops.key_bits(relations[b], c) > 32:
                    raise UnsupportedQuery("join keys >= 2^32: the exchange carries packed runs only")
                joined.add(b)
        if not joined:
            raise UnsupportedQuery("query without predicates")
        for b, _ in selects:
            if b not in joined:
                raise UnsupportedQuery(f"projected binding {b} takes part in no predicate")

        ent: Dict[int, int] = {}          # binding -> row-id column (aligned, this rank's slice)
        carried: Dict[Tuple[int, int], int] = {}  # (binding, column) -> 4-byte key column, aligned with ent
        own_rows: Optional[int] = None    # binding whose ids are rows of this rank's own window

        def window(b):
            return row_window(ops.rows(relations[b]), me, world)

        if filters:
            b0 = fbind[0]
            begin, count = window(b0)
            _, c, op, x = filters[0]
            ids = ops.filter_window(relations[b0], c, op, x, begin, count)
            for _, c, op, x in filters[1:]:
                ids = ops.filter_refine(ids, relations[b0], c, op, x)
            ent[b0] = ids
            own_rows = b0
        for (b, c1), (_, c2) in self_joins:
            if b not in ent or own_rows != b:
                raise UnsupportedQuery("self-join predicate on a binding without a prior filter (reference exits, 8c-iii)")
            new = ops.self_join(relations[b], c1, c2, ent[b])
            ops.free_ids(ent[b])
            ent[b] = new

        remaining = list(chain)
        while remaining:
            pick = None
            for j in remaining:
                ins = [(j[0][0] in ent), (j[1][0] in ent)]
                if not ent or ins[0] != ins[1]:
                    pick = j
                    break
            if pick is None:
                raise UnsupportedQuery("join between two bindings that are both (or neither) in the intermediate result")
            remaining.remove(pick)
            (eb, ec), (nb_, nc) = pick if (not ent or pick[0][0] in ent) else (pick[1], pick[0])
            future = {(b, c) for j in remaining for (b, c) in j}
            key_bits = max(ops.key_bits(relations[eb], ec), ops.key_bits(relations[nb_], nc))
            if key_bits > 32:
                raise UnsupportedQuery("join keys >= 2^32: the exchange carries packed runs only")
            # ---- entity side
            Lcols: List[Tuple[tuple, int]] = []
            temp: List[int] = []
            if not ent or own_rows == eb:
                if not ent:
                    begin, count = window(eb)
                    L = ops.build_window(relations[eb], ec, begin, count)
                    need = sorted(c for (b, c) in future if b == eb)
                    if need:
                        idcol = ops.iota(begin, count, ops.rows(relations[eb]))
                        Lcols.append((("id", eb), idcol))
                        temp.append(idcol)
                        for c in need:
                            h = ops.narrow_window(relations[eb], c, begin, count)
                            Lcols.append((("key", eb, c), h))
                            temp.append(h)
                else:
                    L = ops.build_from_ids(relations[eb], ec, ent[eb])
                    need = sorted(c for (b, c) in future if b == eb)
                    if need:
                        Lcols.append((("id", eb), ent[eb]))
                        for c in need:
                            h = ops.gather_column(relations[eb], c, ent[eb])
                            Lcols.append((("key", eb, c), h))
                            temp.append(h)
            else:
                if (eb, ec) not in carried:
                    raise UnsupportedQuery(f"join key {eb}.{ec} was not carried")
                L = ops.tuples_from_u32(carried[(eb, ec)], key_bits)
                for b, h in ent.items():
                    Lcols.append((("id", b), h))
                for (b, c), h in carried.items():
                    if (b, c) in future:
                        Lcols.append((("key", b, c), h))
            # ---- new base relation
            begin, count = window(nb_)
            R = ops.build_window(relations[nb_], nc, begin, count)
            Rcols: List[Tuple[tuple, int]] = []
            need = sorted(c for (b, c) in future if b == nb_)
            if need:
                idcol = ops.iota(begin, count, ops.rows(relations[nb_]))
                Rcols.append((("id", nb_), idcol))
                temp.append(idcol)
                for c in need:
                    h = ops.narrow_window(relations[nb_], c, begin, count)
                    Rcols.append((("key", nb_, c), h))
                    temp.append(h)
            key_max = max(ops.key_max(relations[eb], ec), ops.key_max(relations[nb_], nc))
            (L2, Lv), (R2, Rv) = self._exchange_pair([(L, [h for _, h in Lcols]), (R, [h for _, h in Rcols])], key_bits, key_max)
            ops.free_tuples(L)
            ops.free_tuples(R)
            for h in temp:
                ops.free_ids(h)
            for h in list(ent.values()) + list(carried.values()):
                ops.free_ids(h)
            ent, carried, own_rows = {}, {}, None
            # ---- local join of this rank's key range
            ops.sort(L2)
            ops.sort(R2)
            pL, pR = ops.merge_join(L2, R2)
            self.stats["local_join_input"] = self.stats.get("local_join_input", 0) + ops.tuples_count(L2) + ops.tuples_count(R2)
            ops.free_tuples(L2)
            ops.free_tuples(R2)
            for names, views, pos, binding in ((Lcols, Lv, pL, eb), (Rcols, Rv, pR, nb_)):
                if not names:
                    ent[binding] = pos       # payloads were row ids
                    continue
                for (name, _), v in zip(names, views):
                    g = ops.gather(v, pos)
                    if name[0] == "id":
                        ent[name[1]] = g
                    else:
                        carried[(name[1], name[2])] = g
                    ops.free_ids(v)
                ops.free_ids(pos)

        if not ent:
            raise UnsupportedQuery("query without predicates")
        nrows = ops.count(next(iter(ent.values())))
        sums = self._project(relations, ent, selects, nrows)
        for h in list(ent.values()) + list(carried.values()):
            ops.free_ids(h)
        vec = np.array([sums[s] for s in selects] + [nrows, self.stats.get("local_join_input", 0)], dtype=np.uint64)
        red = self.comm.allreduce_u64(vec)
        return {"sums": [int(x) for x in red[:len(selects)]], "pairs": int(red[len(selects)]),
                "join_input_tuples": int(red[len(selects) + 1])}


def format_result(res: dict) -> str:
    """The line the reference prints for the same query (print_sums,
    /root/reference/src/utilities.c:212-223)."""
    return "".join("NULL " if res["pairs"] == 0 else f"{s} " for s in res["sums"]) + "\n"


# ---------------------------------------------------------------------------- bench (N > 1)
def bench(eng, host_lib, dist, torch, rank, world, rows, steps, warmup, verify=True):
    """Weak scaling of config 2: every rank OWNS a `rows`-row window of two
    relations of world*rows rows (row-sharded, generated on the device; nothing is
    replicated).  value: windows resident in HBM.  e2e: every step first copies the
    rank's six column windows from pinned host memory and re-derives their statistics."""
    import ctypes as C
    dev = torch.device("cuda", torch.cuda.current_device())
    comm = Comm(dist, torch, dev, rank, world)
    n = world * rows
    begin, count = row_window(n, rank, world)
    assert (begin, count) == (rank * rows, rows)
    gen = torch.Generator(device=dev)
    cols = {}
    for r, seed in enumerate((1, 2)):
        gen.manual_seed(1000 + 64 * seed + rank)
        cols[(r, 0)] = torch.arange(begin, begin + rows, dtype=torch.int64, device=dev)
        cols[(r, 1)] = torch.randint(0, n, (rows,), dtype=torch.int64, device=dev, generator=gen)
        cols[(r, 2)] = torch.randint(0, 10 ** 6, (rows,), dtype=torch.int64, device=dev, generator=gen)
    torch.cuda.synchronize()

    def register():
        for (r, c), t in cols.items():
            mx = comm.allreduce_max(eng.column_max_device(t.data_ptr(), rows))
            eng.adopt_column_window(r, c, t.data_ptr(), begin, rows, n, mx)

    register()
    open_windows(eng, comm, 24 * rows + (64 << 20))
    ex = ShardedExecutor(EngineOps(eng), comm)
    q = "0 1|0.1=1.1&0.2>500000|0.0 1.0 1.2"

    def sync_all():
        eng.sync()
        torch.cuda.synchronize()
        dist.barrier()

    res = ex.run_query(q)
    want = format_result(res)
    verified = None
    if verify and n < (1 << 30):
        # the unsharded host layer (parse -> arrange -> execute_filter/join -> print_sums) on rank 0
        # over replicas of the same columns, once, untimed; the replicas are dropped afterwards
        full = {}
        for (r, c), t in cols.items():
            f = torch.empty(n, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(f, t)
            torch.cuda.synchronize()  # the engine reads f on its own stream (column statistics)
            if rank == 0:
                full[(r, c)] = f
                eng.upload_column_device(2 + r, c, f.data_ptr(), n, adopt=True)
            del f
        torch.cuda.synchronize()
        if rank == 0:
            buf = C.create_string_buffer(4096)
            failed = C.c_int(0)
            host_lib.qce_host_run_batch(b"2 3|0.1=1.1&0.2>500000|0.0 1.0 1.2\n", buf, 4096, C.byref(failed))
            verified = (buf.value.decode() == want) and failed.value == 0
            eng.sync()
        full.clear()
        torch.cuda.empty_cache()
    for _ in range(warmup):
        ex.run_query(q)
    sync_all()
    times, ex_times = [], []
    ncoll0 = comm.n_collectives
    for _ in range(steps):
        sync_all()
        eng.timer_reset()
        res = ex.run_query(q)
        ms, launches = eng.timer_read()
        torch.cuda.synchronize()
        times.append(ms)
        ex_times.append((ex.stats.get("exchange_s", 0.0) + ex.stats.get("project_exchange_s", 0.0)) * 1e3)
        assert format_result(res) == want
    ncoll = (comm.n_collectives - ncoll0) / steps - 1  # minus sync_all's barrier
    t = torch.tensor([sum(times), sum(ex_times)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, ex_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / steps
    sent = torch.tensor([ex.stats["bytes_sent_off_rank"]], dtype=torch.float64, device=dev)
    dist.all_reduce(sent, op=dist.ReduceOp.MAX)

    # per-kernel times of one more pass (CUDA events around every launch; kept out of `value`)
    eng.profile(True)
    for _ in range(2):
        ex.run_query(q)
    prof = {k: v for k, v in eng.profile_read().items() if not k.startswith("gap_")}
    eng.profile(False)
    kernels = {k: round(v["ms"] / 2, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    push_ms = sum(v for k, v in kernels.items() if k.startswith("push_"))
    local_tuples = int(ex.stats.get("local_join_input", 0))  # tuples this rank received and sorted

    # ---- e2e: host windows in, checksums out, every step
    pinned = {k: v.cpu().pin_memory() for k, v in cols.items()}
    e2e_times = []
    for i in range(min(warmup, 1) + steps):
        sync_all()
        eng.timer_reset()
        for k, t in cols.items():
            t.copy_(pinned[k], non_blocking=True)
        torch.cuda.synchronize()
        register()
        res = ex.run_query(q)
        ms, _ = eng.timer_read()
        if i >= min(warmup, 1):
            e2e_times.append(ms)
        assert format_result(res) == want
    te = torch.tensor([sum(e2e_times)], dtype=torch.float64, device=dev)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0]) / steps
    off_rank = float(sent[0])
    return {
        "value": 2.0 * n / (ms_per_step / 1e3), "ms_per_step": ms_per_step,
        "config": {"workload": "C2 weak-scaled: 2-way equi-join + range filter, 2 x %d-row uint64 relations x 3 columns "
                               "(%d rows per GPU per relation, ROW-SHARDED: each GPU holds only its window), query %s, "
                               "sharded by key range over %d GPUs: histogram all-gather, partition+push kernel over "
                               "NVLink peer windows (no all-to-all), row ids pushed to their owners for the checksums, "
                               "uint64 all-reduce" % (n, rows, q, world),
                   "rows_per_relation": n, "result": want.strip(), "verified_against_unsharded_engine": verified,
                   "l2": "inputs larger than L2"},
        "exchange": {"host_ms_per_step_incl_collectives": ex_ms / steps, "push_kernels_ms_per_step": push_ms,
                     "max_bytes_pushed_off_rank_per_step": off_rank,
                     "push_gbs_per_rank": off_rank / (push_ms / 1e3) / 1e9 if push_ms else None,
                     "nvlink_peak_gbs_per_direction": 900.0, "small_collectives_per_step": ncoll},
        "kernels_ms_per_step_rank0": kernels,
        "rank0": {"local_join_input_tuples": local_tuples, "msd_partition_launches_per_step": prof.get("msd_partition", {}).get("launches", 0) // 2},
        "gpu_launches": int(launches) * steps,
        "e2e": {"value": 2.0 * n / (e2e_ms / 1e3), "unit": "rows/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 6 * rows * 8, "d2h_bytes_per_step": 5 * 8,
                "note": "per rank: its own row windows of the six columns from pinned host memory, column statistics "
                        "recomputed, checksums back to the host"},
    }

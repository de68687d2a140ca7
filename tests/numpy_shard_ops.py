"""Numpy stand-in for shardexec.EngineOps (test infrastructure): same op
interface on host arrays, row-sharded columns, and receive "windows" that are
filled through gloo -- every push ships (destination byte offset, bytes) so the
executor's offset bookkeeping is what decides where data lands."""
import numpy as np
import torch.distributed as dist

from oracle import qce_oracle as orc

M32 = np.uint64(0xFFFFFFFF)


def _window(rows, rank, world, align=4096):
    per = -(-rows // world)
    per = -(-per // align) * align
    begin = min(rank * per, rows)
    return begin, min(per, rows - begin)


class NumpyShardOps:
    def __init__(self, db, rank, world, window_bytes=1 << 24):
        self.rank, self.world = rank, world
        self.rows_of = [len(cols[0]) for cols in db]
        self.maxv = [[int(c.max()) if len(c) else 0 for c in cols] for cols in db]
        self.win_of = [_window(n, rank, world) for n in self.rows_of]
        # only this rank's rows are resident: any access outside raises
        self.local = [[c[b:b + n].copy() for c in cols] for cols, (b, n) in zip(db, self.win_of)]
        self.win = np.zeros(window_bytes, dtype=np.uint8)

    # ---- resident rows
    def _col(self, rel, col, ids):
        b, n = self.win_of[rel]
        ids = np.asarray(ids, dtype=np.int64)
        assert ids.size == 0 or (ids.min() >= b and ids.max() < b + n), "row not resident on this rank"
        return self.local[rel][col][ids - b]

    def key_bits(self, rel, col):
        return max(1, self.maxv[rel][col].bit_length())

    def key_max(self, rel, col):
        return self.maxv[rel][col]

    def rows(self, rel, col=0):
        return self.rows_of[rel]

    def filter_window(self, rel, col, op, c, begin, count):
        assert (begin, count) == self.win_of[rel] or count == 0
        return orc.filter_scan(self.local[rel][col][:count], op, c) + np.uint64(begin)

    def filter_refine(self, ids, rel, col, op, c):
        return ids[orc._cmp(self._col(rel, col, ids), op, c)]

    def self_join(self, rel, c1, c2, ids):
        return ids[self._col(rel, c1, ids) == self._col(rel, c2, ids)]

    def build_from_ids(self, rel, col, ids):
        return (self._col(rel, col, ids) << np.uint64(32)) | ids.astype(np.uint64)

    def build_window(self, rel, col, begin, count):
        return (self.local[rel][col][:count] << np.uint64(32)) | np.arange(begin, begin + count, dtype=np.uint64)

    def narrow_window(self, rel, col, begin, count):
        return self.local[rel][col][:count].astype(np.uint32)

    def gather_column(self, rel, col, ids):
        return self._col(rel, col, ids).astype(np.uint32)

    def iota(self, begin, count, bound):
        return np.arange(begin, begin + count, dtype=np.uint32)

    def tuples_from_u32(self, keys, key_bits):
        return (keys.astype(np.uint64) << np.uint64(32)) | np.arange(len(keys), dtype=np.uint64)

    def histogram(self, t, key_bits):
        d = ((t >> np.uint64(32)) >> np.uint64(max(key_bits - 8, 0))) & np.uint64(255)
        return np.bincount(d.astype(np.int64), minlength=256).astype(np.uint64)

    # ---- pushes: (dst, byte offset, payload) delivered through gloo
    def _deliver(self, parcels):
        box = [None] * self.world
        dist.all_gather_object(box, parcels)
        for sender in box:
            for dst, off, raw in sender:
                if dst == self.rank and len(raw):
                    assert off + len(raw) <= len(self.win)
                    self.win[off:off + len(raw)] = np.frombuffer(raw, dtype=np.uint8)

    def push_tuples(self, t, key_bits, splitters, nparts, dst_word_offset, dst_run_index, want_slots):
        part = np.searchsorted(np.array(splitters, dtype=np.uint64), t >> np.uint64(32), side="right") \
            if nparts > 1 else np.zeros(len(t), dtype=np.int64)
        slots = np.zeros(len(t), dtype=np.uint32)
        parcels = []
        for d in range(nparts):
            idx = np.nonzero(part == d)[0]
            seg = t[idx].copy()
            slots[idx] = (np.uint32(d) << np.uint32(28)) | np.arange(len(idx), dtype=np.uint32)
            if dst_run_index is not None:
                seg = (seg & ~M32) | (np.uint64(int(dst_run_index[d])) + np.arange(len(idx), dtype=np.uint64))
            parcels.append((d, int(dst_word_offset[d]) * 8, seg.tobytes()))
        self._deliver(parcels)
        return slots if want_slots else None

    def push_tuples_cols(self, t, key_bits, splitters, nparts, dst_word_offset, dst_run_index, cols, col_u32_offset):
        slots = self.push_tuples(t, key_bits, splitters, nparts, dst_word_offset, dst_run_index, True)
        for c, col in enumerate(cols):
            self.push_col(col, slots, nparts, np.asarray(col_u32_offset)[c])

    def push_col(self, col, slots, nparts, dst_u32_offset):
        parcels = []
        for d in range(nparts):
            idx = np.nonzero((slots >> np.uint32(28)) == d)[0]
            seg = np.zeros(len(idx), dtype=np.uint32)
            seg[(slots[idx] & np.uint32(0x0FFFFFFF)).astype(np.int64)] = col[idx]
            parcels.append((d, int(dst_u32_offset[d]) * 4, seg.tobytes()))
        self._deliver(parcels)

    def _bins(self, ids, per, width, bpr, world):
        ids = np.asarray(ids, dtype=np.int64)
        r = np.minimum(ids // per, world - 1)
        return r * bpr + np.minimum((ids - r * per) // width, bpr - 1)

    def ids_hist(self, ids, per, width, bpr, world):
        return np.bincount(self._bins(ids, per, width, bpr, world), minlength=bpr * world).astype(np.uint64)

    def push_ids(self, ids, per, width, bpr, world, bin_u32_offset):
        bins = self._bins(ids, per, width, bpr, world)
        parcels = []
        for b in np.unique(bins):
            seg = np.asarray(ids)[bins == b].astype(np.uint32)
            parcels.append((int(b) // bpr, int(bin_u32_offset[b]) * 4, seg.tobytes()))
        self._deliver(parcels)

    def fence(self):
        pass

    def tuples_view(self, word_offset, n, key_bits, id_bound, key_range):
        t = self.win[word_offset * 8:(word_offset + n) * 8].view(np.uint64).copy()
        assert n == 0 or ((t >> np.uint64(32)).min() >= key_range[0] and (t >> np.uint64(32)).max() <= key_range[1])
        return t

    def col_view(self, u32_offset, n, id_bound=0, bucketed=False, id_min=0):
        return self.win[u32_offset * 4:(u32_offset + n) * 4].view(np.uint32).copy()

    # ---- local join
    def sort(self, t):
        t[:] = t[np.argsort(t >> np.uint64(32), kind="stable")]

    def merge_join(self, L, R):
        a, b = orc.merge_join(L >> np.uint64(32), L & M32, R >> np.uint64(32), R & M32)
        return a.astype(np.uint32), b.astype(np.uint32)

    def gather(self, col, index):
        return col[np.asarray(index, dtype=np.int64)]

    def checksum(self, ids, rel, cols):
        out = []
        for c in cols:
            v = self._col(rel, c, ids)
            out.append(int(np.add.reduce(v, dtype=np.uint64)) if len(v) else 0)
        return out

    def count(self, ids):
        return len(ids)

    def tuples_count(self, t):
        return len(t)

    def free_ids(self, h):
        pass

    def free_tuples(self, h):
        pass

    def window_bytes(self):
        return len(self.win)

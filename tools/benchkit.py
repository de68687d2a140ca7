"""benchkit.py -- synthetic workloads of the BASELINE configs and their INDEPENDENT checkers.

Bench / test infrastructure (imported by bench.py and tests/), never by the product.

Every column of configs 3, 4 and 5 is a pure function of (relation, column, row index):
  * primary keys are affine permutations  pk(i) = (a * i + b) mod n,  gcd(a, n) = 1;
  * foreign keys are  fk(i) = pk_target(t(i))  with  t(i) = mix(i, seed) mod n_target  the row
    of the target relation they point at (config 4: t = inverse-CDF Zipf(1.2));
  * filter / payload columns are  mix(i, seed) mod domain.
So (a) every rank can generate exactly its row window on its own device, (b) the same
relations can be written to files at a small scale for the CPU reference (the "scaled twin",
SURVEY.md 8d), and (c) the checker needs no join at all: the partner of row i is t(i) by
construction, a chain of joins is a chain of index maps, and a projection checksum is one
vectorised sum over the rows of the driving relation -- linear time, no sort, no merge, no
shared code with the engine.  Arithmetic is uint64 with wrap-around on both sides (torch int64
wraps; numpy uint64).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

M64 = (1 << 64) - 1
DEVICE = "cuda"  # tests on a CPU-only box set this to "cpu"


# ------------------------------------------------------------------ the hash, numpy and torch
def _s64(c: int) -> int:
    """two's-complement view of a 64-bit constant (torch has no uint64 arithmetic)"""
    c &= M64
    return c - (1 << 64) if c >> 63 else c


_C1, _C2, _C3 = 0x9E3779B97F4A7C15, 0xBF58476D1CE4E5B9, 0x94D049BB133111EB


def mix_np(i: np.ndarray, seed: int) -> np.ndarray:
    """splitmix64 finaliser of (i + seed * C1); uint64 in, 63-bit non-negative out."""
    with np.errstate(over="ignore"):
        x = i.astype(np.uint64) + np.uint64((seed * _C1) & M64)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(_C2)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(_C3)
        x = x ^ (x >> np.uint64(31))
    return x >> np.uint64(1)


def mix_t(torch, i, seed: int):
    """the same function on an int64 torch tensor (logical shifts emulated by masking)"""
    def lsr(x, s):
        return (x >> s) & ((1 << (64 - s)) - 1)
    x = i + _s64(seed * _C1)
    x = (x ^ lsr(x, 30)) * _s64(_C2)
    x = (x ^ lsr(x, 27)) * _s64(_C3)
    x = x ^ lsr(x, 31)
    return lsr(x, 1)


class Fn:
    """One column as a function of the row index, evaluated with numpy (xp='np') or torch."""

    def __init__(self, kind: str, **kw):
        self.kind, self.kw = kind, kw

    def __call__(self, xp, i):
        k, kw = self.kind, self.kw
        if k == "id":
            return i
        if k == "hash":  # mix(i, seed) mod domain
            return (mix_np(i, kw["seed"]) % np.uint64(kw["mod"])) if xp == "np" else mix_t(xp, i, kw["seed"]) % kw["mod"]
        if k == "pk":    # affine permutation of [0, n)
            if xp == "np":
                return (i.astype(np.uint64) * np.uint64(kw["a"]) + np.uint64(kw["b"])) % np.uint64(kw["n"])
            return (i * kw["a"] + kw["b"]) % kw["n"]
        if k == "fk":    # pk_target(t(i))
            return kw["pk"](xp, kw["t"](xp, i))
        if k == "either":  # a(i) where sel(i) < cut, else b(i)
            s = kw["sel"](xp, i)
            a, b = kw["a"](xp, i), kw["b"](xp, i)
            return np.where(s < np.uint64(kw["cut"]), a, b) if xp == "np" else xp.where(s < kw["cut"], a, b)
        if k == "zipf":  # inverse CDF of Zipf(s) over [0, n): rank 0 is the heavy key
            cdf = kw["cdf"](xp)
            if xp == "np":
                u = (mix_np(i, kw["seed"]) >> np.uint64(10)).astype(np.float64) / float(1 << 53)
                return np.minimum(np.searchsorted(cdf, u), kw["n"] - 1).astype(np.uint64)
            u = (mix_t(xp, i, kw["seed"]) >> 10).to(xp.float64) / float(1 << 53)
            return xp.clamp(xp.searchsorted(cdf, u), max=kw["n"] - 1)
        raise ValueError(k)


def pk_fn(n: int, seed: int) -> Fn:
    a = (0x2545F491 + 2 * seed) | 1  # odd, < 2^31: a * i stays below 2^63 for i < 2^32
    while math.gcd(a, n) != 1:
        a += 2
    return Fn("pk", a=a, b=(seed * 7919 + 13) % max(n, 1), n=max(n, 1))


def row_fn(n_target: int, seed: int) -> Fn:
    return Fn("hash", seed=seed, mod=max(n_target, 1))


class ZipfCdf:
    """cumulative Zipf(s) weights over n ranks, built once per device"""

    def __init__(self, n: int, s: float):
        self.n, self.s, self._np, self._t = n, s, None, None

    def __call__(self, xp):
        if xp == "np":
            if self._np is None:
                w = np.arange(1, self.n + 1, dtype=np.float64) ** (-self.s)
                self._np = np.cumsum(w / w.sum())
            return self._np
        if self._t is None:
            w = xp.arange(1, self.n + 1, dtype=xp.float64, device=DEVICE) ** (-self.s)
            self._t = xp.cumsum(w / w.sum(), 0)
        return self._t


# ------------------------------------------------------------------ workloads
class Workload:
    """relations[r] = (rows, [Fn per column]); queries = text lines; driving info for the checker."""

    def __init__(self, name: str, relations, queries: List[str], describe: str):
        self.name, self.relations, self.queries, self.describe = name, relations, queries, describe

    def rows(self, r: int) -> int:
        return self.relations[r][0]

    def column_np(self, r: int, c: int, begin: int = 0, count: Optional[int] = None) -> np.ndarray:
        n = self.rows(r)
        count = n - begin if count is None else count
        return np.ascontiguousarray(self.relations[r][1][c]("np", np.arange(begin, begin + count, dtype=np.uint64)), dtype=np.uint64)

    def column_t(self, torch, r: int, c: int, begin: int, count: int):
        i = torch.arange(begin, begin + count, dtype=torch.int64, device=DEVICE)
        return self.relations[r][1][c](torch, i).contiguous()

    def text(self, rel_offset: int = 0, first: Optional[int] = None) -> str:
        """the batch as stdin text, relation ids shifted by rel_offset (several workloads share one engine)"""
        out = []
        for q in self.queries[:first]:
            rels, rest = q.split("|", 1)
            out.append(" ".join(str(int(x) + rel_offset) for x in rels.split()) + "|" + rest + "\n")
        return "".join(out)

    def referenced(self) -> List[Tuple[int, int]]:
        """(relation, column) pairs the query batch reads, in first-use order"""
        seen, out = set(), []
        for q in self.queries:
            rels, preds, sels = q.strip().split("|")
            rel = [int(x) for x in rels.split()]
            toks = []
            for p in preds.split("&"):
                for side in p.replace("<", "=").replace(">", "=").split("="):
                    if "." in side:
                        toks.append(side)
            toks += sels.split()
            for t in toks:
                b, c = t.split(".")
                key = (rel[int(b)], int(c))
                if key not in seen:
                    seen.add(key)
                    out.append(key)
        return out

    def join_input_rows(self) -> int:
        """the metric's numerator: base rows of all bindings that enter a join, over the batch"""
        total = 0
        for q in self.queries:
            rels, preds, _ = q.strip().split("|")
            rel = [int(x) for x in rels.split()]
            joined = set()
            for p in preds.split("&"):
                if "=" in p:
                    l, r = p.split("=")
                    if "." in l and "." in r:
                        joined.add(int(l.split(".")[0]))
                        joined.add(int(r.split(".")[0]))
            total += sum(self.rows(rel[b]) for b in joined)
        return total


def c3_workload(rows: int, seed: int = 3, thr: int = 900) -> Workload:
    """Config 3: 4-way PK-FK chain with a self-join predicate and a filter."""
    rel = []
    pks = [pk_fn(rows, seed * 100 + k) for k in range(4)]
    ts = [row_fn(rows, seed * 100 + 10 + k) for k in range(4)]
    for k in range(4):
        fk = Fn("fk", pk=pks[(k + 1) % 4], t=ts[k])
        other = Fn("fk", pk=pks[(k + 1) % 4], t=row_fn(rows, seed * 100 + 20 + k))
        c2 = Fn("either", sel=Fn("hash", seed=seed * 100 + 30 + k, mod=100), cut=50, a=fk, b=other)
        c3 = Fn("hash", seed=seed * 100 + 40 + k, mod=1000)
        rel.append((rows, [pks[k], fk, c2, c3]))
    q = "0 1 2 3|0.1=0.2&0.1=1.0&1.1=2.0&2.1=3.0&0.3<%d|0.3 1.3 2.3 3.3" % thr
    w = Workload("c3", rel, [q], "4-way join chain with filter and self-join predicate, %d rows per relation" % rows)
    w.ts, w.thr = ts, thr
    return w


def c3_check(torch, w: Workload, begin: int, count: int):
    """partial (pairs, sums[4]) of the rows [begin, begin+count) of relation 0: index maps only"""
    i = torch.arange(begin, begin + count, dtype=torch.int64, device=DEVICE)
    c = w.relations[0][1]
    keep = (c[1](torch, i) == c[2](torch, i)) & (c[3](torch, i) < w.thr)
    i0 = i[keep]
    i1 = w.ts[0](torch, i0)
    i2 = w.ts[1](torch, i1)
    i3 = w.ts[2](torch, i2)
    sums = [int(w.relations[k][1][3](torch, ix).sum()) & M64 for k, ix in enumerate((i0, i1, i2, i3))]
    return int(i0.numel()), sums


def c4_workload(rows: int, seed: int = 4, s: float = 1.2) -> Workload:
    """Config 4: 3-way join, FK side Zipf(1.2)-skewed against a unique PK."""
    pk1, pk2 = pk_fn(rows, seed * 100 + 1), pk_fn(rows, seed * 100 + 2)
    z = Fn("zipf", cdf=ZipfCdf(rows, s), seed=seed * 100 + 3, n=rows)
    t1 = row_fn(rows, seed * 100 + 4)
    r0 = (rows, [Fn("id"), Fn("fk", pk=pk1, t=z), Fn("hash", seed=seed * 100 + 5, mod=1000)])
    r1 = (rows, [pk1, Fn("fk", pk=pk2, t=t1), Fn("hash", seed=seed * 100 + 6, mod=1000)])
    r2 = (rows, [pk2, Fn("hash", seed=seed * 100 + 7, mod=1000), Fn("hash", seed=seed * 100 + 8, mod=1000)])
    w = Workload("c4", [r0, r1, r2], ["0 1 2|0.1=1.0&1.1=2.0|0.0 1.2 2.1"],
                 "skewed-key (Zipf %.1f) 3-way join, %d rows per relation" % (s, rows))
    w.z, w.t1 = z, t1
    return w


def c4_check(torch, w: Workload, begin: int, count: int):
    i0 = torch.arange(begin, begin + count, dtype=torch.int64, device=DEVICE)
    i1 = w.z(torch, i0)
    i2 = w.t1(torch, i1)
    sums = [int(i0.sum()) & M64, int(w.relations[1][1][2](torch, i1).sum()) & M64,
            int(w.relations[2][1][1](torch, i2).sum()) & M64]
    return int(i0.numel()), sums


C5_NREL = 14


def c5_sizes(scale: float = 1.0) -> List[int]:
    """14 relation sizes, log-uniform 10^6 .. 10^9 rows (deterministic: evenly spaced in the log)"""
    return [max(64, int(round(10 ** (6 + 3 * k / (C5_NREL - 1)) * scale))) for k in range(C5_NREL)]


def c5_workload(scale: float = 1.0, nqueries: int = 1000, seed: int = 5) -> Workload:
    """Config 5: 14 relations (4 columns: c0 = id, c1 = FK -> relation r+1, c2 = FK -> relation r+5,
    c3 = uniform[0,1000)), ~1000 mixed queries: filter-only, and 1..3-join FK->PK chains with a
    filter on the first binding (inside the parity-defined query class)."""
    sizes = c5_sizes(scale)
    rel = []
    ts = {}
    for r, n in enumerate(sizes):
        t1, t2 = row_fn(sizes[(r + 1) % C5_NREL], seed * 1000 + r * 4 + 1), row_fn(sizes[(r + 5) % C5_NREL], seed * 1000 + r * 4 + 2)
        ts[(r, 1)], ts[(r, 2)] = t1, t2
        rel.append((n, [Fn("id"), t1, t2, Fn("hash", seed=seed * 1000 + r * 4 + 3, mod=1000)]))
    rng = np.random.default_rng(seed)
    queries, plans = [], []
    for _ in range(nqueries):
        kind = rng.choice(4, p=[0.25, 0.35, 0.25, 0.15])  # joins in the query
        r0 = int(rng.integers(0, C5_NREL))
        thr = int(rng.integers(20, 700))
        chain = [r0]
        hops = []
        for _j in range(kind):
            c = int(rng.choice([1, 2]))
            hops.append(c)
            chain.append((chain[-1] + (1 if c == 1 else 5)) % C5_NREL)
        if len(set(chain)) != len(chain):  # the same relation twice: outside the parity class (8c-v)
            chain, hops = chain[:1], []
        preds = ["%d.%d=%d.0" % (k, hops[k], k + 1) for k in range(len(hops))] + ["0.3<%d" % thr]
        if not hops:
            preds = ["0.3<%d" % thr, "0.1>%d" % int(rng.integers(0, max(2, sizes[(r0 + 1) % C5_NREL] // 2)))]
        sel = ["%d.%d" % (b, int(rng.choice([0, 3]))) for b in range(len(chain))]
        if not hops:
            sel = ["0.0", "0.2"]
        queries.append(" ".join(map(str, chain)) + "|" + "&".join(preds) + "|" + " ".join(sel))
        plans.append((chain, hops, thr, preds, sel))
    w = Workload("c5", rel, queries, "%d mixed queries (0-3 FK->PK joins + filters) over 14 relations of %d..%d rows"
                 % (nqueries, sizes[0], sizes[-1]))
    w.ts, w.plans = ts, plans
    return w


def c5_check_query(torch, w: Workload, k: int, begin: int, count: int):
    """partial result of query k over rows [begin, begin+count) of its first relation.
    Returns (count_line or None, pairs, sums)."""
    chain, hops, thr, preds, sel = w.plans[k]
    i = torch.arange(begin, begin + count, dtype=torch.int64, device=DEVICE)
    c = w.relations[chain[0]][1]
    keep = c[3](torch, i) < thr
    first = None
    if not hops:
        # two stacked filters: the second one executed refines the first one's row ids and the
        # reference prints how many survive (src/filter.c:32) -- the rows that pass both
        lim = int(preds[1].split(">")[1])
        keep = keep & (c[1](torch, i) > lim)
        first = int(keep.sum())
    idx = [i[keep]]
    for h, r in zip(hops, chain[:-1]):
        idx.append(w.ts[(r, h)](torch, idx[-1]))
    sums = []
    for s in sel:
        b, col = (int(x) for x in s.split("."))
        sums.append(int(w.relations[chain[b]][1][col](torch, idx[b]).sum()) & M64)
    return first, int(idx[0].numel()), sums


def format_line(pairs: int, sums: Sequence[int]) -> str:
    """print_sums' line (/root/reference/src/utilities.c:212-223)"""
    return "".join("NULL " if pairs == 0 else "%d " % s for s in sums) + "\n"


# ------------------------------------------------------------------ config 2 (numpy default_rng family of round 1)
def c2_check(torch, dist, c1_l, c2_l, c0_l, c1_r, c0_r, c2_r, key_domain: int, thr: int):
    """Independent closed form for `0 1|0.1=1.1&0.2>thr|0.0 1.0 1.2` over (this rank's windows of)
    the six columns: multiplicities by bincount over the key domain, all-reduced when sharded.
    Returns (lhs, pairs, sums[3])."""
    keep = c2_l > thr
    kl = c1_l[keep]
    cnt_l = torch.bincount(kl, minlength=key_domain)
    cnt_r = torch.bincount(c1_r, minlength=key_domain)
    if dist is not None:
        dist.all_reduce(cnt_l)
        dist.all_reduce(cnt_r)
    lhs = int(keep.sum())
    s0 = int((c0_l[keep] * cnt_r[kl]).sum()) & M64
    wl_ = cnt_l[c1_r]
    s1 = int((c0_r * wl_).sum()) & M64
    s2 = int((c2_r * wl_).sum()) & M64
    pairs = int(wl_.sum())
    out = [lhs, pairs, s0, s1, s2]
    if dist is not None:
        t = torch.tensor([x - (1 << 64) if x >> 63 else x for x in out], dtype=torch.int64, device=DEVICE)
        dist.all_reduce(t)
        out = [int(x) & M64 for x in t.tolist()]
    return out[0], out[1], out[2:]

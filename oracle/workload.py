"""oracle/workload.py -- TEST INFRASTRUCTURE ONLY.

Seeded synthetic relations / query batches in the reference's formats, a runner
for the compiled reference binary (oracle/_ref/queries), the relational-truth
evaluator and the three-way PDQ classifier of SURVEY.md 8c.

File format (fill_data, /root/reference/src/utilities.c:105-121): uint64 tuple
count, uint64 column count, then the columns one after another (column-major
uint64).  stdin protocol (read_relations :124-162, parser src/parsing.c:118-148):
relation paths one per line, `Done`, then `rels|preds|selects` lines to EOF.
"""
from __future__ import annotations

import os
import subprocess
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import qce_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
REF_BIN = os.path.join(HERE, "_ref", "queries")
ALTRAND = os.path.join(HERE, "_ref", "altrand.so")
U64 = np.uint64


# --------------------------------------------------------------------------- files
def write_relation(path: str, cols: Sequence[np.ndarray]) -> None:
    n = len(cols[0])
    with open(path, "wb") as f:
        np.array([n, len(cols)], dtype=U64).tofile(f)
        for c in cols:
            assert len(c) == n
            np.ascontiguousarray(c, dtype=U64).tofile(f)


def read_relation(path: str) -> List[np.ndarray]:
    raw = np.fromfile(path, dtype=U64)
    n, c = int(raw[0]), int(raw[1])
    return [raw[2 + j * n: 2 + (j + 1) * n] for j in range(c)]


def write_db(dirname: str, db: Sequence[Sequence[np.ndarray]]) -> List[str]:
    os.makedirs(dirname, exist_ok=True)
    paths = []
    for i, cols in enumerate(db):
        p = os.path.join(dirname, f"r{i}")
        write_relation(p, cols)
        paths.append(p)
    return paths


def stdin_text(paths: Sequence[str], queries: str) -> str:
    return "".join(p + "\n" for p in paths) + "Done\n" + queries


def have_reference() -> bool:
    return os.access(REF_BIN, os.X_OK)


def run_binary(binary: str, paths: Sequence[str], queries: str, altrand: bool = False,
               timeout: float = 600.0, env_extra: Optional[Dict[str, str]] = None):
    env = dict(os.environ)
    if altrand:
        env["LD_PRELOAD"] = ALTRAND
    if env_extra:
        env.update(env_extra)
    p = subprocess.run([binary], input=stdin_text(paths, queries).encode(), stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, env=env, timeout=timeout)
    return p.stdout.decode(), p.stderr.decode(), p.returncode


def run_reference(paths, queries, altrand=False, timeout=600.0):
    return run_binary(REF_BIN, paths, queries, altrand=altrand, timeout=timeout)


# --------------------------------------------------------------------------- truth
def truth_query(q: orc.Query, db) -> str:
    """Relational semantics of one query (connected join graphs only): the
    checksum line a *correct* engine prints.  Independent of the reference's
    state machine; used to classify reference outputs (PDQ-T vs PDQ-D)."""
    nb = len(q.relations)
    sel = {}
    for p in q.predicates:
        if p.type == 1:
            b = p.first[0]
            col = db[q.relations[b]][p.first[1]]
            ids = sel.get(b)
            if ids is None:
                ids = np.arange(len(col), dtype=np.int64)
            sel[b] = ids[orc._cmp(col[ids], p.op, p.second)]
    table: Dict[int, np.ndarray] = {}
    pending = [p for p in q.predicates if p.type == 0]

    def base(b):
        ids = sel.get(b)
        return ids if ids is not None else np.arange(len(db[q.relations[b]][0]), dtype=np.int64)

    while pending:
        progressed = False
        for p in list(pending):
            (lb, lc), (rb, rc) = p.first, p.second
            lcol, rcol = db[q.relations[lb]][lc], db[q.relations[rb]][rc]
            if not table:
                if lb == rb:
                    ids = base(lb)
                    table[lb] = ids[lcol[ids] == rcol[ids]]
                    pending.remove(p); progressed = True
                    continue
                table[lb] = base(lb)
            if lb in table and rb in table:
                keep = lcol[table[lb]] == rcol[table[rb]]
                for k in table:
                    table[k] = table[k][keep]
            elif lb in table or rb in table:
                (ib, icol), (ob, ocol) = ((lb, lcol), (rb, rcol)) if lb in table else ((rb, rcol), (lb, lcol))
                oids = base(ob)
                okeys = ocol[oids]
                order = np.argsort(okeys, kind="stable")
                okeys, oids = okeys[order], oids[order]
                ikeys = icol[table[ib]]
                lo = np.searchsorted(okeys, ikeys, "left")
                hi = np.searchsorted(okeys, ikeys, "right")
                cnt = hi - lo
                tot = int(cnt.sum())
                rep = np.repeat(np.arange(len(ikeys)), cnt)
                offs = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
                for k in table:
                    table[k] = table[k][rep]
                table[ob] = oids[np.repeat(lo, cnt) + offs]
            else:
                continue
            pending.remove(p)
            progressed = True
        if not progressed:
            raise ValueError("disconnected join graph")
    out = []
    for b, c in q.selects:
        if table:
            if b not in table:
                raise ValueError("select on a binding outside the join graph")
            ids = table[b]
        elif b in sel:
            ids = sel[b]
        else:
            raise ValueError("select on a binding without any predicate")
        col = db[q.relations[b]][c]
        with np.errstate(over="ignore"):
            out.append("NULL " if len(ids) == 0 else f"{int(np.sum(col[ids], dtype=U64))} ")
    return "".join(out) + "\n"


def classify(paths, db, query_line: str, timeout=120.0) -> Tuple[str, str]:
    """Three runs of SURVEY 8c: reference, reference with alt rand(), truth.
    Returns (class, reference_stdout); class in
    'PDQ-T' (tie-invariant and == truth), 'PDQ-D' (tie-invariant, != truth),
    'TIE' (differs under alt rand), 'CRASH' (non-zero exit / signal)."""
    try:
        o1, e1, rc1 = run_reference(paths, query_line, timeout=timeout)
        o2, e2, rc2 = run_reference(paths, query_line, altrand=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return "CRASH", ""
    if rc1 != 0 or rc2 != 0:
        return "CRASH", o1
    if o1 != o2:
        return "TIE", o1
    q = orc.parse_query(query_line)
    try:
        t = truth_query(q, db)
    except ValueError:
        return "PDQ-D", o1
    last = o1.splitlines(keepends=True)[-1] if o1 else ""
    return ("PDQ-T" if last == t else "PDQ-D"), o1


# --------------------------------------------------------------------------- generators
def gen_small_db(seed: int = 2018, scale: float = 1.0):
    """C1: 'SIGMOD'18-small-shaped' database: 14 relations, 1K-1.5M
    rows at scale 1 (use scale << 1 in unit tests), 3-9 columns, col0 = unique id,
    the others uniform in small domains or foreign keys into another relation's
    col0."""
    rng = np.random.default_rng(seed)
    sizes = [max(8, int(s * scale)) for s in
             [1561, 3754, 4598, 2000, 12000, 60000, 150000, 300000, 500000, 800000, 1000000, 1200000, 1500000, 40000]]
    ncols = [3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 9, 4, 3]
    db = []
    for r, (n, c) in enumerate(zip(sizes, ncols)):
        cols = [np.arange(n, dtype=U64)]
        for j in range(1, c):
            kind = (r + j) % 3
            if kind == 0:  # FK into another relation's id column
                tgt = sizes[(r + j) % len(sizes)]
                cols.append(rng.integers(0, tgt, n, dtype=np.uint64))
            elif kind == 1:  # small domain
                cols.append(rng.integers(0, 1000, n, dtype=np.uint64))
            else:  # medium domain
                cols.append(rng.integers(0, max(2, n // 2), n, dtype=np.uint64))
        db.append(cols)
    return db


def gen_pair_db(n: int, domain: int, seeds=(1, 2), filt_domain: int = 10 ** 6):
    """C2-shaped: two relations, 3 uint64 columns: c0 = i, c1 = uniform[0,domain),
    c2 = uniform[0,filt_domain) (SURVEY 8d worked example)."""
    db = []
    for s in seeds:
        rng = np.random.default_rng(s)
        db.append([np.arange(n, dtype=U64), rng.integers(0, domain, n, dtype=np.uint64),
                   rng.integers(0, filt_domain, n, dtype=np.uint64)])
    return db


def gen_chain_db(n: int, nrel: int = 4, seed: int = 3, self_frac: float = 0.5):
    """C3-shaped PK-FK chain: col0 = random permutation PK, col1 = uniform FK into
    the next relation's PK, col2 = col1 on a fraction of rows (self-join predicate
    0.1=0.2), col3 = uniform[0,1000)."""
    rng = np.random.default_rng(seed)
    db = []
    for _ in range(nrel):
        pk = rng.permutation(n).astype(U64)
        fk = rng.integers(0, n, n, dtype=np.uint64)
        c2 = np.where(rng.random(n) < self_frac, fk, rng.integers(0, n, n, dtype=np.uint64)).astype(U64)
        c3 = rng.integers(0, 1000, n, dtype=np.uint64)
        db.append([pk, fk, c2, c3])
    return db


def gen_zipf_db(n: int, seed: int = 4, s: float = 1.2):
    """C4-shaped: R0.c1 ~ Zipf(s) over R1's PK values (FK-side skew only), R1.c1
    uniform FK into R2's PK."""
    rng = np.random.default_rng(seed)
    ranks = np.arange(1, n + 1, dtype=np.float64)
    w = ranks ** (-s)
    cdf = np.cumsum(w / w.sum())
    z = np.searchsorted(cdf, rng.random(n)).clip(0, n - 1).astype(U64)
    perm = rng.permutation(n).astype(U64)
    r0 = [np.arange(n, dtype=U64), perm[z], rng.integers(0, 1000, n, dtype=np.uint64)]
    r1 = [rng.permutation(n).astype(U64), rng.integers(0, n, n, dtype=np.uint64), rng.integers(0, 1000, n, dtype=np.uint64)]
    r2 = [rng.permutation(n).astype(U64), rng.integers(0, 1000, n, dtype=np.uint64), rng.integers(0, 1000, n, dtype=np.uint64)]
    return [r0, r1, r2]


def gen_queries(db, count: int, seed: int = 7, max_joins: int = 3) -> List[str]:
    """Random query lines in the template mix of SURVEY 8d/C1: 1..max_joins join
    predicates over distinct relations forming a chain/star, optional filters on
    ONE binding, the filter never written first.  Candidates only -- admit them to
    a parity batch through classify()."""
    rng = np.random.default_rng(seed)
    out = []
    nrel = len(db)
    while len(out) < count:
        nj = int(rng.integers(0, max_joins + 1))
        nb = nj + 1
        rels = [int(x) for x in rng.choice(nrel, size=nb, replace=False)]
        preds = []
        for j in range(nj):
            a = int(rng.integers(0, j + 1))
            b = j + 1
            ca = int(rng.integers(0, len(db[rels[a]])))
            cb = int(rng.integers(0, len(db[rels[b]])))
            preds.append(f"{a}.{ca}={b}.{cb}")
        nf = int(rng.integers(0, 3)) if nj else int(rng.integers(1, 3))
        fb = int(rng.integers(0, nb))
        filts = []
        for _ in range(nf):
            c = int(rng.integers(0, len(db[rels[fb]])))
            col = db[rels[fb]][c]
            v = int(col[int(rng.integers(0, len(col)))])
            op = "<>="[int(rng.integers(0, 3))]
            filts.append(f"{fb}.{c}{op}{min(v, 2**32 - 1)}")
        if nj:
            allp = preds[:1] + filts + preds[1:]
        else:
            allp = filts
        ns = int(rng.integers(1, 4))
        sels = []
        for _ in range(ns):
            b = int(rng.integers(0, nb))
            if nj == 0:
                b = fb
            sels.append(f"{b}.{int(rng.integers(0, len(db[rels[b]])))}")
        out.append(" ".join(map(str, rels)) + "|" + "&".join(allp) + "|" + " ".join(sels))
    return out

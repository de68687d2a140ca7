"""Regenerates tests/golden/*.json|npz by RUNNING THE REFERENCE (oracle/_ref,
built from /root/reference by oracle/Makefile).  Run in the build container:

    make -C oracle && python tests/golden/make_golden.py

The reference ships no golden vectors of its own (no tests/, workload files
git-ignored), so these fixtures are its recorded outputs:
  arrange.json      predicate order after the reference's arrange_predicates
  small_db.npz      a scaled C1-shaped database (14 relations)
  small_batch.json  candidate queries with the reference's stdout and their
                    PDQ class (reference / reference+alt-rand / relational truth)
  ops_db.npz, ops.json   the operational-behaviour cases of SURVEY.md 8a'
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import workload as wl  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PROBE = os.path.join(ROOT, "oracle", "_ref", "arrange_probe")


def save_db(path, db):
    np.savez_compressed(path, **{f"r{r}_c{c}": col for r, cols in enumerate(db) for c, col in enumerate(cols)})


def golden_arrange():
    rng = np.random.default_rng(42)
    preds = [
        "0.1=1.1&1.2=2.1&0.2<100", "0.1=1.1&0.2<100&1.2=2.1&2.2>5", "0.1=1.0&1.1=2.0&2.1=3.0&0.1=0.2&0.3<500",
        "0.1=0.2&0.1=1.0&1.1=2.0&2.1=3.0&0.3<900", "1.1=2.1&0.1=1.1&0.2=2.2", "0.1=1.1&2.1=3.1&1.1=2.2&0.1=3.2",
        "0.2=1.0&0.1=2.0&0.2>3499", "0.2<100&0.1=1.1&1.1=2.1", "0.1<10&0.2>500", "0.1=1.1", "0.2>7",
        "0.2<100&1.1=2.1&0.1=1.1&2.2>4", "0.0=1.0&0.1<5&0.2<6&0.3<7",
    ]
    for _ in range(60):
        n = int(rng.integers(1, 6))
        toks = []
        for _ in range(n):
            if rng.random() < 0.3:
                toks.append(f"{rng.integers(0,4)}.{rng.integers(0,4)}{'<>='[rng.integers(0,3)]}{rng.integers(0,5000)}")
            else:
                toks.append(f"{rng.integers(0,4)}.{rng.integers(0,4)}={rng.integers(0,4)}.{rng.integers(0,4)}")
        preds.append("&".join(toks))
    text = "".join(f"0 1 2 3|{p}|0.0\n" for p in preds)
    out = subprocess.run([PROBE], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout.decode().splitlines()
    assert len(out) == len(preds)
    json.dump([{"written": p, "executed": o} for p, o in zip(preds, out)],
              open(os.path.join(HERE, "arrange.json"), "w"), indent=0)
    print("arrange.json:", len(preds))


def classify_batch(db, queries, name):
    d = tempfile.mkdtemp()
    paths = wl.write_db(d, db)
    recs, hist = [], {}
    for q in queries:
        cls, ref = wl.classify(paths, db, q + "\n", timeout=60)
        recs.append({"query": q, "class": cls, "stdout": ref})
        hist[cls] = hist.get(cls, 0) + 1
    json.dump(recs, open(os.path.join(HERE, name), "w"), indent=0)
    print(name, hist)


def golden_small():
    db = wl.gen_small_db(seed=2018, scale=0.004)
    save_db(os.path.join(HERE, "small_db.npz"), db)
    qs = wl.gen_queries(db, 160, seed=7, max_joins=3)
    classify_batch(db, qs, "small_batch.json")


def golden_ops():
    """SURVEY 8a' operational behaviours, on tiny relations."""
    rng = np.random.default_rng(1)
    sizes = [2000, 2000, 500, 800]
    db = []
    for n in sizes:
        db.append([np.arange(n, dtype=np.uint64), rng.integers(0, 700, n, dtype=np.uint64),
                   rng.integers(0, 1000, n, dtype=np.uint64), rng.integers(0, 1000, n, dtype=np.uint64)])
    # PK-FK chain material: r_k.c1 is a FK into r_{k+1}.c0 (unique)
    for k in range(3):
        db[k][1] = rng.integers(0, sizes[k + 1], sizes[k], dtype=np.uint64)
    save_db(os.path.join(HERE, "ops_db.npz"), db)
    qs = [
        "0 1|0.2>5000|0.0",                       # filter matches nothing -> NULL
        "0|0.1<10&0.2>500|0.0 0.2",               # stacked filters: count line first
        "0|0.2<500&0.1=0.2|0.0",                  # self-join after a filter on the binding
        "0|0.1=0.2|0.0",                          # self-join on a fresh binding: exit(1)
        "0 1|0.1=1.1&0.2<20&1.2<20|0.0 1.0",      # two filtered bindings: positional scan join
        "0 0|0.1=1.1&0.2<20&1.2<20|0.0 1.0",      # same relation, same column: join skipped
        "0 1|0.1=1.0|0.0 1.0 0.2 1.3",            # plain PK-FK join, 4 checksums
        "0 1|0.1=1.0&0.2<300|0.2 1.2",            # filter + join
        "0 1|0.1=1.0&0.2<300&0.3>100|0.2 1.2",    # two filters on one binding + join
        "0 1 2|0.1=1.0&1.1=2.0|0.0 1.0 2.0",      # 2 joins chain
        "0 1 2|0.1=1.0&1.1=2.0&0.3<500|0.3 1.3 2.3",
        "0 1 2 3|0.1=1.0&1.1=2.0&2.1=3.0&0.3<500|0.3 1.3 2.3 3.3",   # 3-join PK-FK chain
        "0 1 2 3|0.1=1.1&1.2=2.1&2.2=3.1|0.0 1.0 2.0 3.0",           # tie-dependent 3-join
        "0 1|0.2=1.2|0.0 1.0",                    # many-to-many on a small domain
        "0 1|0.2=1.2&0.1=1.1|0.0 1.0",            # second predicate on joined bindings -> scan join
        "1 0|0.1=1.1&1.2>900|0.0 1.0",            # filter on the rhs binding
        "0 1|0.1=1.0&1.0=0.1|0.0",                # repeated predicate
        "0 1 2|0.1=1.0&0.1=2.0|0.0 1.0 2.0",      # star on one column: JOIN_SORT path
    ]
    classify_batch(db, qs, "ops.json")


def golden_edge():
    """Empty / one-row / duplicate-heavy / tile-boundary relations, repeated bindings."""
    rng = np.random.default_rng(5)
    db = [
        [np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint64)],                                   # r0: no rows
        [np.array([7], dtype=np.uint64), np.array([3], dtype=np.uint64), np.array([9], dtype=np.uint64)],  # r1: one row
        [np.array([1, 1, 1, 2, 2], dtype=np.uint64), np.array([3, 3, 3, 3, 7], dtype=np.uint64),
         np.array([5, 6, 7, 8, 9], dtype=np.uint64)],                                                   # r2: duplicates
        [np.arange(4097, dtype=np.uint64), rng.integers(0, 8, 4097, dtype=np.uint64)],                  # r3: 4097 rows
        [rng.integers(0, 8, 4096, dtype=np.uint64), np.arange(4096, dtype=np.uint64),
         rng.integers(0, 3, 4096, dtype=np.uint64)],                                                    # r4: 4096 rows
    ]
    save_db(os.path.join(HERE, "edge_db.npz"), db)
    qs = [
        "0|0.0<5|0.0", "0 1|0.0=1.0|0.1 1.1", "1 0|0.0=1.0|0.1", "1|0.0=7|0.0 0.1 0.2", "1|0.0>7|0.0", "1|0.1<4&0.2>8|0.2",
        "1 2|0.1=1.1|0.0 1.2", "2 2|0.0=1.1|0.2 1.2", "2 2|0.1=1.1|0.2 1.2", "2 2|0.0=1.0&0.2<8|0.2 1.2",
        "2 3|0.1=1.1|0.2 1.0", "3 4|0.1=1.0|0.0 1.1", "3 4|0.1=1.0&0.0<100|0.0 1.1 1.2", "3 4|0.1=1.0&1.2=1|0.0 1.1",
        "4 3|0.0=1.1&0.2=0&0.1>4000|0.1 1.0", "3 4 2|0.1=1.0&1.2=2.0|0.0 1.1 2.2", "3|0.1=3|0.0 0.1", "3|0.0>4095|0.0",
        "3|0.0<1&0.1<100|0.1", "4 4|0.2=1.2&0.1<40&0.0=5|0.1 1.1", "3 3|0.0=1.0|0.1 1.1", "2|0.0=0.1|0.2", "2|0.1<5&0.0=0.1|0.2",
    ]
    classify_batch(db, qs, "edge.json")


if __name__ == "__main__":
    assert wl.have_reference(), "build oracle/_ref first: make -C oracle"
    golden_arrange()
    golden_ops()
    golden_small()
    golden_edge()

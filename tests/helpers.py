"""Shared test helpers: golden fixtures, database files, binaries."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
PKG = os.path.join(ROOT, "query-compiler-executor_b200")
QUERIES_BIN = os.path.join(PKG, "build", "queries")
REFMAIN_BIN = os.path.join(PKG, "build", "queries_refmain")
HOST_PROBE = os.path.join(PKG, "build", "host_probe")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_db(name):
    z = np.load(os.path.join(GOLDEN, name))
    rels = {}
    for key in z.files:
        r, c = key[1:].split("_c")
        rels.setdefault(int(r), {})[int(c)] = z[key]
    return [[rels[r][c] for c in sorted(rels[r])] for r in sorted(rels)]


def load_json(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def run_queries_bin(binary, paths, query_text, timeout=300, env=None):
    text = "".join(p + "\n" for p in paths) + "Done\n" + query_text
    full_env = dict(os.environ)
    if env:
        full_env.update({k: str(v) for k, v in env.items()})
    p = subprocess.run([binary], input=text.encode(), stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout,
                       env=full_env)
    return p.stdout.decode(), p.stderr.decode(), p.returncode


def choose_splitters(hist, key_bits, nparts):
    """numpy restatement of the splitter choice inside qce_exchange_plan: splitter keys on
    histogram-bin boundaries so that every part gets about total/nparts tuples.  hist = global
    256-bin histogram of the top 8 significant key bits.  part(key) = #splitters <= key."""
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


def row_window(rows, rank, world, align=4096):
    """numpy restatement of qce_row_share"""
    per = -(-rows // world)
    per = -(-per // align) * align
    begin = min(rank * per, rows)
    return begin, min(per, rows - begin)

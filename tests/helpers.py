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

import os, sys, tempfile
sys.path.insert(0, os.getcwd())
from tests.helpers import QUERIES_BIN, load_db, load_json, run_queries_bin
from oracle import workload as wl
db = load_db("small_db.npz")
paths = wl.write_db(tempfile.mkdtemp(), db)
recs = [r for r in load_json("small_batch.json") if r["class"] in ("PDQ-T", "PDQ-D")]
env = {"QCE_REPLICATE_BYTES": 0, "QCE_COMM_TIMEOUT_S": 30, "QCE_GPUS": 2}
bad = 0
for i, r in enumerate(recs):
    o, e, rc = run_queries_bin(QUERIES_BIN, paths, r["query"] + "\n", env=env, timeout=120)
    refused = o == "" and "refused" in e
    if not (o == r["stdout"] or (r["class"] == "PDQ-D" and refused)):
        bad += 1
        print("### MISMATCH", i, r["class"], r["query"], "rc", rc, "\n got ", repr(o), "\n want", repr(r["stdout"]), "\n err:", e[:1500])
        if bad >= 6:
            break
print("checked", i + 1, "bad", bad)

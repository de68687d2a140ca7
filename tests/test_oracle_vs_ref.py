"""CPU: the numpy oracle against the live reference binary (oracle/_ref, only
where /root/reference was available to build it) on fresh random queries."""
import tempfile

import pytest

from oracle import qce_oracle as orc
from oracle import workload as wl

pytestmark = pytest.mark.skipif(not wl.have_reference(), reason="oracle/_ref/queries not built")


def test_random_queries_against_live_reference():
    db = wl.gen_small_db(seed=99, scale=0.002)
    paths = wl.write_db(tempfile.mkdtemp(), db)
    n_checked = 0
    for q in wl.gen_queries(db, 30, seed=123):
        cls, ref = wl.classify(paths, db, q + "\n", timeout=60)
        if cls in ("PDQ-T", "PDQ-D"):
            assert orc.run_batch(db, q + "\n") == ref, q
            n_checked += 1
    assert n_checked >= 20


def test_c2_scaled_twin_against_live_reference():
    db = wl.gen_pair_db(20000, 20000, filt_domain=1000)
    paths = wl.write_db(tempfile.mkdtemp(), db)
    q = "0 1|0.1=1.1&0.2>500|0.0 1.0 1.2\n"
    ref, _, rc = wl.run_reference(paths, q)
    assert rc == 0 and orc.run_batch(db, q) == ref

"""CPU: the ranks' shared-memory communicator inside libqce_b200.so (csrc/qce_comm.cuh) -- the
all-gather / all-reduce / gather-to-root the sharded operators and the batch scheduler agree
through.  Pure host code: runs without a GPU, three real processes."""
import ctypes as C
import multiprocessing as mp
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _rank_main(rank, world, name, token, q, fail_rank):
    import qce_b200
    lib = qce_b200.load_library()
    try:
        assert lib.qce_comm_attach(name.encode(), rank, world, token) == 0, lib.qce_last_error()
        assert lib.qce_comm_rank() == rank and lib.qce_comm_world() == world
        v = np.array([rank + 1, 10 * (rank + 1), (1 << 64) - 1], dtype=np.uint64)
        assert lib.qce_comm_allreduce_sum_u64(v.ctypes.data, 3) == 0
        m = np.array([rank, 7], dtype=np.uint64)
        assert lib.qce_comm_allreduce_max_u64(m.ctypes.data, 2) == 0
        blob = (b"rank%d;" % rank) * (rank + 1) * (40000 if rank == 1 else 1)  # > one 1 MB slot for rank 1
        out = C.c_void_p()
        lens = (C.c_uint64 * world)()
        assert lib.qce_comm_gatherv(blob, len(blob), C.byref(out), lens) == 0
        gathered = C.string_at(out.value, sum(lens)) if rank == 0 else b""
        for _ in range(200):
            assert lib.qce_comm_barrier() == 0
        res = {"sum": v.tolist(), "max": m.tolist(), "lens": list(lens), "gathered_ok": None}
        if rank == 0:
            want = b"".join((b"rank%d;" % r) * (r + 1) * (40000 if r == 1 else 1) for r in range(world))
            res["gathered_ok"] = gathered == want
        if fail_rank is not None:
            # one rank fails: the others' next wait must end with an error, not hang
            if rank == fail_rank:
                lib.qce_comm_abort()
                res["after_abort"] = "raised"
            else:
                res["after_abort"] = lib.qce_comm_barrier()
        q.put((rank, res))
    except Exception as e:  # noqa: BLE001
        q.put((rank, {"error": repr(e)}))


def _run(world, fail_rank=None):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name, token = "qce_test_%d_%d" % (os.getpid(), world), 0x1234567 + world
    ps = [ctx.Process(target=_rank_main, args=(r, world, name, token, q, fail_rank)) for r in range(world)]
    for p in ps:
        p.start()
    got = dict(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=30)
    return got


def test_collectives_three_ranks():
    got = _run(3)
    for r in range(3):
        assert "error" not in got[r], got[r]
        assert got[r]["sum"] == [6, 60, ((1 << 64) - 1) * 3 % (1 << 64)]
        assert got[r]["max"] == [2, 7]
        assert got[r]["lens"] == [6, 6 * 2 * 40000, 6 * 3]
    assert got[0]["gathered_ok"] is True


def test_abort_reaches_every_rank():
    got = _run(2, fail_rank=1)
    assert "error" not in got[0], got[0]
    assert got[0]["after_abort"] == -1


def test_single_rank_needs_no_segment():
    import qce_b200
    lib = qce_b200.load_library()
    v = np.array([5], dtype=np.uint64)
    assert lib.qce_comm_world() == 1 and lib.qce_comm_rank() == 0
    assert lib.qce_comm_allreduce_sum_u64(v.ctypes.data, 1) == 0 and v[0] == 5
    assert lib.qce_comm_barrier() == 0

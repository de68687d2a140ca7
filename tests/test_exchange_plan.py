"""CPU: the native exchange bookkeeping (qce_exchange_plan / qce_rowid_push_plan: host
arithmetic inside libqce_b200.so, no device needed) against an independent numpy
restatement, on random histograms.  Every rank must derive the same layout, and the
segments of one destination must tile its window without overlap."""
import numpy as np
import pytest

import qce_b200
from qce_b200 import engine as eng_mod
from tests.helpers import choose_splitters


def _numpy_plan(H, world, rank, ncols, key_bits):
    nsides = len(ncols)
    splitters = choose_splitters(H.sum(axis=(0, 1)), key_bits, world)
    shift = max(key_bits - 8, 0)
    bnd = np.minimum(np.array([0] + [sp >> shift for sp in splitters] + [256], dtype=np.int64), 256)
    Hc = np.zeros((world, nsides, 257), dtype=np.int64)
    np.cumsum(H.astype(np.int64), axis=2, out=Hc[:, :, 1:])
    C = (Hc[:, :, bnd[1:]] - Hc[:, :, bnd[:-1]]).transpose(1, 0, 2)
    recv = C.sum(axis=1)
    before = (np.cumsum(C, axis=1) - C)[:, rank, :]
    run_off, col_off = [], []
    top = np.zeros(world, dtype=np.int64)
    for k in range(nsides):
        run_off.append(top)
        top = (top + 8 * recv[k] + 15) // 16 * 16
        for _ in range(ncols[k]):
            col_off.append(top)
            top = (top + 4 * recv[k] + 15) // 16 * 16
    sent = np.array([C[k, rank].sum() - C[k, rank, rank] for k in range(nsides)])
    return splitters, recv, before, np.array(run_off), np.array(col_off).reshape(-1, world), int(top.max()), sent, C


@pytest.mark.parametrize("world", [1, 2, 3, 8, 16])
@pytest.mark.parametrize("key_bits", [5, 8, 20, 28, 32])
def test_exchange_plan_matches_numpy(world, key_bits):
    lib = qce_b200.load_library()
    rng = np.random.default_rng(world * 100 + key_bits)
    for trial in range(6):
        nsides = int(rng.integers(1, 4))
        ncols = [int(x) for x in rng.integers(0, 4, nsides)]
        H = rng.integers(0, 50_000, (world, nsides, 256)).astype(np.uint64)
        if trial == 1:
            H[:, :, 10:] = 0          # everything in a few bins: later parts are empty
        if trial == 2:
            H[:] = 0                   # empty runs
        if trial == 3:
            H[:, :, 7] += np.uint64(10 ** 7)   # one heavy bin
        layouts = []
        for rank in range(world):
            sp, recv, before, run_off, col_off, need, sent = eng_mod.exchange_plan(lib, H, world, rank, ncols, key_bits)
            esp, erecv, ebefore, erun, ecol, eneed, esent, C = _numpy_plan(H, world, rank, ncols, key_bits)
            assert sp == esp
            np.testing.assert_array_equal(recv.astype(np.int64), erecv)
            np.testing.assert_array_equal(before.astype(np.int64), ebefore)
            np.testing.assert_array_equal(run_off.astype(np.int64), erun)
            if sum(ncols):
                np.testing.assert_array_equal(col_off.astype(np.int64), ecol)
            assert need == eneed
            np.testing.assert_array_equal(sent.astype(np.int64), esent)
            layouts.append((sp, recv.tobytes(), run_off.tobytes()))
        assert all(l == layouts[0] for l in layouts)   # every rank derives the same plan
        # the segments [before, before + C) of all ranks tile each destination's run exactly
        for k in range(nsides):
            for d in range(world):
                starts = [int(eng_mod.exchange_plan(lib, H, world, r, ncols, key_bits)[2][k, d]) for r in range(world)]
                sizes = [int(C[k, r, d]) for r in range(world)]
                assert starts == list(np.cumsum([0] + sizes[:-1]))
                assert sum(sizes) == int(erecv[k, d])


def _numpy_rowid_plan(H, world, rank, bpr):
    nbind, nb = H.shape[1], bpr * world
    owner = np.arange(nb) // bpr
    first_bin = owner * bpr
    top = np.zeros(world, dtype=np.int64)
    offs, views, sent = [], [], []
    for k in range(nbind):
        Hk = H[:, k, :].astype(np.int64)
        bin_tot = Hk.sum(axis=0)
        cs = np.cumsum(bin_tot) - bin_tot
        bin_start = cs - cs[first_bin]
        total = np.add.reduceat(bin_tot, np.arange(0, nb, bpr))
        src_before = Hk[:rank].sum(axis=0)
        region = top
        top = (top + 4 * total + 15) // 16 * 16
        offs.append(region[owner] // 4 + bin_start + src_before)
        views.append((int(region[rank]) // 4, int(total[rank])))
        sent.append(int(Hk[rank].sum() - Hk[rank, rank * bpr:(rank + 1) * bpr].sum()))
    return np.array(offs), views, int(top.max()), sent


@pytest.mark.parametrize("world,bpr", [(1, 1), (2, 1), (3, 1), (8, 1), (2, 128), (3, 85), (8, 32), (16, 1)])
def test_rowid_push_plan_matches_numpy(world, bpr):
    lib = qce_b200.load_library()
    rng = np.random.default_rng(world * 7 + bpr)
    for nbind in (1, 2, 3):
        H = rng.integers(0, 100_000, (world, nbind, bpr * world)).astype(np.uint64)
        for rank in range(world):
            offs, voff, vcnt, need, sent = eng_mod.rowid_push_plan(lib, H, world, rank, nbind, bpr)
            eoffs, eviews, eneed, esent = _numpy_rowid_plan(H, world, rank, bpr)
            np.testing.assert_array_equal(offs.astype(np.int64), eoffs)
            assert [(int(a), int(b)) for a, b in zip(voff, vcnt)] == eviews
            assert need == eneed and [int(x) for x in sent] == esent

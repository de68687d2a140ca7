"""GPU parity of the peer-memory exchange primitives (k_exchange.cuh) on ONE GPU:
the local window is presented as `world` ranks (qce_xwin_loopback), so every
store of the push kernels lands where a multi-GPU run would put it, and the
sharded operators themselves are exercised by tests/test_gpu_ranks.py (real ranks, same kernels)
against the relational truth / the oracle."""
import numpy as np
import pytest

from oracle import qce_oracle as orc
from oracle import workload as wl

pytestmark = pytest.mark.gpu
U64 = np.uint64
WIN = 64 << 20


@pytest.fixture(scope="module")
def xeng(engine):
    engine.xwin_create(WIN)
    yield engine
    engine.xwin_destroy()


def _read_words(e, byte_off, n, key_bits):
    t = e.tuples_from_window(byte_off // 8, n, key_bits)
    k, p = e.tuples_to_host(t)
    e.tuples_free(t)
    return k, p


def _read_u32(e, byte_off, n):
    h = e.rowids_from_window(byte_off // 4, n)
    v = e.rowids_to_host(h)
    e.rowids_free(h)
    return v


@pytest.mark.parametrize("n", [0, 1, 33, 4095, 4096, 4097, 100003, 1 << 20])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("rewrite", [False, True, "fused"])
def test_push_tuples_loopback(xeng, n, world, rewrite):
    e = xeng
    e.xwin_loopback(world)
    per = WIN // world // 4096 * 4096
    rng = np.random.default_rng(n + world)
    key_bits = 20
    keys = rng.integers(0, 1 << key_bits, n, dtype=np.uint64)
    ids = rng.permutation(n).astype(U64)
    t = e.tuples_from_host(keys, ids)
    hist = e.key_histogram(t, key_bits)
    from tests.helpers import choose_splitters
    splitters = choose_splitters(hist, key_bits, world)
    part = np.searchsorted(np.array(splitters, dtype=U64), keys, side="right") if world > 1 else np.zeros(n, dtype=np.int64)
    counts = np.bincount(part, minlength=world)
    seg_words = np.array([64 + 2 * d for d in range(world)], dtype=U64)     # where "this rank's" segment starts
    run_index = np.array([1000 * (d + 1) for d in range(world)], dtype=np.uint32)
    col = rng.integers(0, 1 << 32, n, dtype=np.uint64)
    col_off = np.array([(per // 2) // 4 + 16 * d for d in range(world)], dtype=U64)
    fused = rewrite == "fused"
    if fused:   # tuples and two bystander columns in ONE kernel; the second column is col ^ 5
        ch, ch2 = e.rowids_from_host(col), e.rowids_from_host(col ^ U64(5))
        col2_off = col_off + U64(per // 16)
        e.push_tuples_cols(t, key_bits, splitters, world, seg_words, run_index, [ch, ch2], np.stack([col_off, col2_off]))
        e.rowids_free(ch)
        e.rowids_free(ch2)
    else:
        slots_h = e.push_tuples(t, key_bits, splitters, world, seg_words, run_index if rewrite else None, bool(rewrite))
    if rewrite and not fused:
        ch = e.rowids_from_host(col)
        e.push_u32_by_slot(ch, slots_h, world, col_off)
        slots = e.rowids_to_host(slots_h)
        e.rowids_free(ch)
        e.rowids_free(slots_h)
        assert np.array_equal(slots >> U64(28), part.astype(U64))
    e.sync()
    for d in range(world):
        k, p = _read_words(e, per * d + int(seg_words[d]) * 8, int(counts[d]), key_bits)
        want = part == d
        if not rewrite:
            got = np.lexsort((p, k))
            exp = np.lexsort((ids[want], keys[want]))
            np.testing.assert_array_equal(k[got], keys[want][exp])
            np.testing.assert_array_equal(p[got], ids[want][exp])
        else:
            # payload = index in the receiver's run; the bystander column sits at the same index
            np.testing.assert_array_equal(np.sort(p), run_index[d] + np.arange(counts[d], dtype=U64))
            got_col = _read_u32(e, per * d + int(col_off[d]) * 4, int(counts[d]))
            pos = (p - run_index[d]).astype(np.int64)
            pairs_got = sorted(zip(k.tolist(), got_col[pos].tolist()))
            pairs_exp = sorted(zip(keys[want].tolist(), col[want].tolist()))
            assert pairs_got == pairs_exp
            if fused:
                got2 = _read_u32(e, per * d + int(col2_off[d]) * 4, int(counts[d]))
                np.testing.assert_array_equal(got2, got_col ^ U64(5))
                continue
            # slots say where each input tuple went
            idx = np.nonzero(want)[0]
            np.testing.assert_array_equal(k[np.argsort(pos)][(slots[idx] & U64(0x0FFFFFFF)).astype(np.int64)], keys[idx])


@pytest.mark.parametrize("n", [0, 5, 4097, 8193, 300001])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("fine", [False, True])
def test_push_rowids_loopback(xeng, n, world, fine):
    e = xeng
    e.xwin_loopback(world)
    per_bytes = WIN // world // 4096 * 4096
    rng = np.random.default_rng(7 * n + world)
    rows = 1_000_003
    rows_per_rank = -(-(-(-rows // world)) // 4096) * 4096
    bpr = 256 // world if fine else 1  # one bin per owner (ballot-ranked kernel) or L2-sized regions
    width = -(-rows_per_rank // bpr)
    ids = rng.integers(0, rows, n, dtype=np.uint64)
    h = e.rowids_from_host(ids)
    hist = e.rowids_bin_histogram(h, rows_per_rank, width, bpr, world)
    r = np.minimum(ids // U64(rows_per_rank), U64(world - 1))
    bins = (r * U64(bpr) + np.minimum((ids - r * U64(rows_per_rank)) // U64(width), U64(bpr - 1))).astype(np.int64)
    np.testing.assert_array_equal(hist, np.bincount(bins, minlength=bpr * world).astype(U64))
    # bin-major layout inside each owner, 8 ids of slack before every bin (a stand-in for other senders' segments)
    offs = np.zeros(bpr * world, dtype=U64)
    for d in range(world):
        at = 32
        for b in range(d * bpr, (d + 1) * bpr):
            offs[b] = at + 8
            at += 8 + int(hist[b])
    e.push_rowids(h, rows_per_rank, width, bpr, world, offs)
    e.sync()
    for b in range(bpr * world):
        got = _read_u32(e, per_bytes * (b // bpr) + int(offs[b]) * 4, int(hist[b]))
        np.testing.assert_array_equal(np.sort(got), np.sort(ids[bins == b]))
    e.rowids_free(h)


def test_window_overflow_and_misuse_are_reported(xeng):
    import qce_b200
    e = xeng
    e.xwin_loopback(2)
    with pytest.raises(qce_b200.EngineError, match="exceeds the exchange window"):
        e.tuples_from_window(WIN // 8, 16, 20)
    t = e.tuples_from_host(np.arange(10, dtype=U64), np.arange(10, dtype=U64))
    with pytest.raises(qce_b200.EngineError, match="differs from the attached world size"):
        e.push_tuples(t, 8, [128, 192], 3, np.zeros(3, dtype=U64))
    with pytest.raises(qce_b200.EngineError, match="not on a boundary"):
        e.push_tuples(t, 20, [5], 2, np.zeros(2, dtype=U64))
    e.tuples_free(t)


def test_gather_and_carried_columns(engine):
    rng = np.random.default_rng(3)
    n = 100_003
    col = rng.integers(0, 1 << 30, n, dtype=np.uint64)
    engine.upload_column(120, 0, col)
    w = engine.column_window_u32(120, 0, 4096, 50_000)
    np.testing.assert_array_equal(engine.rowids_to_host(w), col[4096:54096])
    ids = rng.integers(0, n, 70_001, dtype=np.uint64)
    hi = engine.rowids_from_host(ids)
    g = engine.column_gather_u32(120, 0, hi)
    np.testing.assert_array_equal(engine.rowids_to_host(g), col[ids])
    idx = rng.integers(0, 50_000, 33_333, dtype=np.uint64)
    hx = engine.rowids_from_host(idx)
    gg = engine.rowids_gather(w, hx)
    np.testing.assert_array_equal(engine.rowids_to_host(gg), col[4096:54096][idx])
    it = engine.rowids_iota(4096, 1000, n)
    np.testing.assert_array_equal(engine.rowids_to_host(it), np.arange(4096, 5096, dtype=U64))
    t = engine.tuples_from_u32(w, 30)
    k, p = engine.tuples_to_host(t)
    np.testing.assert_array_equal(k, col[4096:54096])
    np.testing.assert_array_equal(p, np.arange(50_000, dtype=U64))
    engine.tuples_free(t)
    for h in (w, hi, g, hx, gg, it):
        engine.rowids_free(h)
    import qce_b200
    engine.upload_column(120, 1, np.array([1 << 40, 3], dtype=U64))
    with pytest.raises(qce_b200.EngineError, match="cannot be carried"):
        engine.column_window_u32(120, 1, 0, 2)


def test_windowed_column_refuses_non_resident_rows(engine):
    import qce_b200
    import torch
    v = torch.arange(8192, dtype=torch.int64, device="cuda")
    engine.adopt_column_window(121, 0, v.data_ptr(), 4096, 8192, 20000, 20000)
    h = engine.filter_scan(121, 0, "<", 6, rows=(4096, 8192))   # values are 0..8191 at rows 4096..12287
    np.testing.assert_array_equal(engine.rowids_to_host(h), np.arange(4096, 4102, dtype=U64))
    engine.rowids_free(h)
    with pytest.raises(qce_b200.EngineError, match="not resident"):
        engine.filter_scan(121, 0, "<", 6)
    with pytest.raises(qce_b200.EngineError, match="not resident"):
        engine.build_tuples(121, 0, rows=(0, 4096))


# ---------------------------------------------------------------- executor, world size 1


def test_failed_ipc_attach_does_not_poison_the_next_launch(xeng):
    """A non-sticky CUDA failure (an IPC handle that cannot be opened) must not leave its error code in
    the runtime's last-error slot: the next kernel launch checks cudaGetLastError() and would report
    the stale failure as its own (CK() now clears it)."""
    import qce_b200
    e = xeng
    with pytest.raises(qce_b200.EngineError):
        e.xwin_attach(2, 0, bytes(128))  # rank 1's "handle" is garbage
    e.xwin_loopback(1)
    col = np.arange(50_000, dtype=U64)
    e.upload_column(61, 0, col)
    ids = e.filter_scan(61, 0, "<", 1000)   # launches kernels right after the failed call
    assert e.rowids_count(ids) == 1000
    e.rowids_free(ids)

"""Single-GPU timing of the push kernels with every store local (loopback window):
separates kernel structure from NVLink effects."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import qce_b200
from qce_b200 import sharded
eng = qce_b200.Engine(0)
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
eng.xwin_create(4 << 30)
dev = torch.device("cuda", 0)
ids_t = torch.randint(0, 2 * n, (n,), dtype=torch.int64, device=dev)
col = torch.randint(0, 1 << 27, (n,), dtype=torch.int64, device=dev)
torch.cuda.synchronize()
eng.upload_column_device(0, 0, col.data_ptr(), n, adopt=True)
t = eng.build_tuples(0, 0)
h = eng.rowids_from_host(ids_t.cpu().numpy().astype(np.uint64))
for world in (2, 8):
    eng.xwin_loopback(world)
    per_bytes = (4 << 30) // world // 4096 * 4096
    key_bits = 27
    hist = eng.key_histogram(t, key_bits)
    spl = sharded.choose_splitters(hist, key_bits, world)
    rows_per_rank = -(-2 * n // world)
    eng.profile(True)
    for _ in range(5):
        eng.push_tuples(t, key_bits, spl, world, np.zeros(world, dtype=np.uint64))
        hh = eng.rowids_bin_histogram(h, rows_per_rank, rows_per_rank, 1, world)
        eng.push_rowids(h, rows_per_rank, rows_per_rank, 1, world, np.full(world, 1 << 20, dtype=np.uint64))
        ph = eng.rowids_bin_histogram(h, rows_per_rank, -(-rows_per_rank // (256 // world)), 256 // world, world)
    eng.sync()
    prof = eng.profile_read()
    eng.profile(False)
    print("world", world, {k: round(v["ms"] / v["launches"], 4) for k, v in prof.items() if not k.startswith("gap")})
    print("   tuples GB/s", 16 * n / (prof["push_tuples"]["ms"] / 5 / 1e3) / 1e9, "ids GB/s", 8 * n / (prof["push_rowids"]["ms"] / 5 / 1e3) / 1e9)

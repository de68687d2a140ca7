"""Scratch probe: config-2 pipeline driven primitive by primitive through the
C-ABI, with per-kernel device times (CUDA events).  Not the bench -- a first
look at where the time goes."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import qce_b200

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
e = qce_b200.Engine()
t0 = time.time()
for r, seed in enumerate((1, 2)):
    rng = np.random.default_rng(seed)
    e.upload_column(r, 0, np.arange(n, dtype=np.uint64))
    e.upload_column(r, 1, rng.integers(0, n, n, dtype=np.uint64))
    e.upload_column(r, 2, rng.integers(0, 10**6, n, dtype=np.uint64))
print("load s", time.time() - t0, flush=True)

def run():
    f = e.filter_scan(0, 2, ">", 500000)
    L = e.build_tuples(0, 1, f)
    R = e.build_tuples(1, 1)
    e.sort_tuples(L); e.sort_tuples(R)
    oL, oR = e.merge_join(L, R)
    s0 = e.checksum(oL, 0, [0]); s1 = e.checksum(oR, 1, [0, 2])
    m = e.rowids_count(oL)
    for h in (f, oL, oR): e.rowids_free(h)
    e.tuples_free(L); e.tuples_free(R)
    return m, s0 + s1

for i in range(reps):
    e.timer_reset()
    w = time.time()
    m, sums = run()
    ms, launches = e.timer_read()
    print(f"rep {i}: pairs {m} sums {sums} device {ms:.3f} ms wall {1e3*(time.time()-w):.3f} ms launches {launches}", flush=True)
e.profile(True)
run()
prof = {k: v for k, v in e.profile_read().items() if not k.startswith("gap_before")}
e.profile(False)
tot = sum(v["ms"] for v in prof.values())
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:16s} {v['launches']:4d} launches {v['ms']:9.3f} ms  {100*v['ms']/tot:5.1f}%")
print("sum of kernels", tot, "ms; input rows/s", 2 * n / (tot / 1e3))
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"n": n, "profile": prof, "pairs": m}, open("gpurun_out/c2_probe.json", "w"))

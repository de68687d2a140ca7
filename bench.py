#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200.

Metric (BASELINE.json): join input rows/s.  Workload at N=1: configs[1] --
single 2-way equi-join + range filter over 2 x 100M-row relations of 3 uint64
columns (c0 = i, c1 = uniform[0, rows), c2 = uniform[0, 1e6)), query
`0 1|0.1=1.1&0.2>500000|0.0 1.0 1.2` (SURVEY.md 8d worked example).  One "step"
= that query through the host operator layer (parse -> arrange -> execute_filter
-> execute_join -> print_sums; libqce_host.so -> libqce_b200.so).

  value   rows/s with the base columns already resident in HBM
  e2e     same, but every step first copies the six referenced columns from
          pinned host memory to the device and reads the checksums back
  roofline  dominant kernel (the MSD partition pass of the sort on these inputs), live CUDA-event times
  cpu_baseline  the reference's own binary (oracle/_ref/queries) on a bounded
          scaled twin of the workload, on this box's host cores

`--impl reference` times that CPU binary as the reference arm.
N>1 (torchrun): the join is sharded by key range across ranks,
weak scaling -- every rank owns a 2 x 100M-row window (row-sharded columns) and
the exchange is a partition kernel that stores into the peers' windows over
NVLink (query-compiler-executor_b200/shardexec.py).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "query-compiler-executor_b200")
METRIC = "join_input_rows_per_s"
QUERY = "0 1|0.1=1.1&0.2>{thr}|0.0 1.0 1.2\n"
REF_SAMPLE_ROWS = 1_000_000  # per relation: ~2-3 s of the (output-quadratic) CPU reference per step


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region.  In-process
    NVML (a `nvidia-smi -lms 100` child perturbed the CUDA driver enough to cost
    ~25 % of a 10 ms step); nvidia-smi at 200 ms is the fallback."""
    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period=0.02):
        self.index, self.period, self.sm, self.reasons, self.max_mhz = index, period, [], set(), None
        self._stop, self._thr, self._smi = threading.Event(), None, None

    def start(self):
        if self.period <= 0:
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

            def loop():
                while not self._stop.is_set():
                    try:
                        self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                        for name, bit in self.BAD.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    self._stop.wait(self.period)
            self._thr = threading.Thread(target=loop, daemon=True)
            self._thr.start()
        except Exception:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            try:
                self._smi = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}",
                                              "--format=csv,noheader,nounits", "-lms", "200"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            except Exception:
                self._smi = None

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1.0)
        if self._smi:
            time.sleep(0.25)
            self._smi.terminate()
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for line in self._smi.stdout.read().splitlines():
                f = [x.strip() for x in line.split(",")]
                try:
                    self.sm.append(float(f[0])); self.max_mhz = float(f[1])
                except Exception:
                    continue
                for nme, v in zip(names, f[2:6]):
                    if v == "Active":
                        self.reasons.add(nme)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


# ------------------------------------------------------------------ workload
def gen_relation(rows, seed, key_domain, row_base=0):
    rng = np.random.default_rng(seed)
    return [np.arange(row_base, row_base + rows, dtype=np.uint64),
            rng.integers(0, key_domain, rows, dtype=np.uint64),
            rng.integers(0, 10 ** 6, rows, dtype=np.uint64)]


def host_lib():
    C.CDLL(os.path.join(PKG, "libqce_b200.so"), mode=C.RTLD_GLOBAL)
    lib = C.CDLL(os.path.join(PKG, "libqce_host.so"))
    lib.qce_host_run_batch.restype = C.c_long
    lib.qce_host_run_batch.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
    return lib


def run_query(lib, text):
    buf = C.create_string_buffer(4096)
    failed = C.c_int(0)
    n = lib.qce_host_run_batch(text.encode(), buf, 4096, C.byref(failed))
    if n < 0 or failed.value:
        raise RuntimeError("host layer failed the query")
    return buf.value.decode()


# ------------------------------------------------------------------ CPU reference arm
def time_reference(rows, steps, warmup):
    """The unmodified reference binary (single-threaded C) on a scaled twin of the
    workload; per step: whole-process wall time minus a load-only run."""
    from oracle import workload as wl
    if not wl.have_reference():
        return None
    db = [gen_relation(rows, 1, rows), gen_relation(rows, 2, rows)]
    d = tempfile.mkdtemp(prefix="qce_ref_")
    paths = wl.write_db(d, db)
    q = QUERY.format(thr=500000)
    t = time.time(); wl.run_reference(paths, ""); load_s = time.time() - t
    times, out = [], ""
    for i in range(warmup + steps):
        t = time.time()
        out, err, rc = wl.run_reference(paths, q)
        dt = time.time() - t
        if rc != 0:
            raise RuntimeError("reference failed: " + err)
        if i >= warmup:
            times.append(max(dt - load_s, 1e-9))
    for p in paths:
        os.unlink(p)
    sec = float(np.mean(times))
    return {"rows_per_s": 2 * rows / sec, "sec_per_step": sec, "load_s": load_s, "stdout": out.strip()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000, help="rows per relation per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nproc = os.cpu_count() or 1

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps, warm = max(1, min(args.steps, 20)), max(0, min(args.warmup, 3))
        r = time_reference(REF_SAMPLE_ROWS, steps, warm)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/queries not built"}))
            return 0
        line = {
            "impl": "reference", "metric": METRIC, "value": r["rows_per_s"], "unit": "rows/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * r["sec_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "C2 scaled twin: 2-way equi-join + range filter, 2 x %d-row uint64 relations, "
                                   "3 checksums (the CPU reference is quadratic in join output, full size is infeasible)"
                                   % REF_SAMPLE_ROWS, "rows_per_relation": REF_SAMPLE_ROWS},
            "cpu_baseline": {"value": r["rows_per_s"], "unit": "rows/s", "cores": 1, "kind": "reference",
                             "sample": "2 x %d rows, process wall time minus load-only run (%.2f s), host has %d cores, "
                                       "reference is single-threaded" % (REF_SAMPLE_ROWS, r["load_s"], nproc)},
            "e2e": {"value": r["rows_per_s"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    import qce_b200
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = qce_b200.Engine(local_rank)
    lib = host_lib()
    rows = args.rows
    pk, pk_src = peaks()

    if world > 1:
        from qce_b200 import sharded, shardexec
        rows = rows // 4096 * 4096  # equal, vector-aligned row windows on every rank
        sampler = ClockSampler(local_rank, period=float(os.environ.get('QCE_BENCH_CLOCK_PERIOD', '0.05')))
        sampler.start()
        if os.environ.get("QCE_EXCHANGE", "push") == "nccl":
            # earlier transport, kept for comparison: replicated columns, send buffer + NCCL all-to-all
            res = sharded.bench(eng, lib, dist, rank, world, rows, args.steps, args.warmup)
        else:
            try:
                res = shardexec.bench(eng, lib, dist, torch, rank, world, rows, args.steps, args.warmup)
            except shardexec.PeerWindowsUnavailable as e:
                # no P2P / CUDA IPC between these GPUs: every rank lands here together
                res = sharded.bench(eng, lib, dist, rank, world, rows, args.steps, args.warmup)
                res["config"]["transport_note"] = "peer windows unavailable (%s): NCCL all-to-all transport" % (e,)
        res["clocks"] = sampler.stop()
        if rank == 0:
            # roofline of the dominant HBM kernel on rank 0 (the MSD partition of the received runs:
            # 8 B read + 8 B written per tuple per level) and the link roofline of the push kernels
            k = res.get("kernels_ms_per_step_rank0", {})
            r0 = res.get("rank0", {})
            if k.get("msd_partition") and r0.get("local_join_input_tuples"):
                levels = max(1, r0.get("msd_partition_launches_per_step", 4) // 2)
                alg = 16.0 * r0["local_join_input_tuples"] * levels
                ach = alg / (k["msd_partition"] / 1e3) / 1e9
                res["roofline"] = {"bound": "hbm", "kernel": "msd_partition (rank 0)", "achieved": ach, "peak": pk["hbm_gbs"],
                                   "peak_source": pk_src, "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": None,
                                   "algorithmic_bytes": "16 B per received tuple per level, %d levels" % levels}
            ex_ = res.get("exchange", {})
            if ex_.get("push_gbs_per_rank"):
                res["nvlink"] = {"bound": "nvlink", "kernel": "push_tuples + push_rowids", "achieved": ex_["push_gbs_per_rank"],
                                 "peak": 900.0, "peak_source": "NVLink 5 nominal per direction per GPU", "unit": "GB/s",
                                 "frac": ex_["push_gbs_per_rank"] / 900.0,
                                 "note": "off-rank bytes / time of the whole push kernels (their local share included in the time)"}
            res_line = res
            res_line.update({"metric": METRIC, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
                             "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                             "dtype": "u64", "data": "synthetic"})
            print(json.dumps(res_line))
        dist.barrier()
        dist.destroy_process_group()
        return 0

    # ---- data: pinned host copies of the six referenced columns
    t0 = time.time()
    host_cols = {}
    for r, seed in enumerate((1, 2)):
        for c, col in enumerate(gen_relation(rows, seed, rows)):
            t = torch.from_numpy(col.view(np.int64)).pin_memory()
            host_cols[(r, c)] = t
    gen_s = time.time() - t0
    col_bytes = rows * 8

    def upload_all():
        for (r, c), t in host_cols.items():
            eng.lib.qce_upload_column(r, c, t.data_ptr(), rows)

    q = QUERY.format(thr=500000)
    upload_all()
    want = run_query(lib, q)  # first (cold) run doubles as the result every later step must reproduce

    sampler = ClockSampler(local_rank, period=float(os.environ.get('QCE_BENCH_CLOCK_PERIOD', '0.05')))
    # ---- value: columns resident in HBM
    for _ in range(args.warmup):
        assert run_query(lib, q) == want
    eng.sync(); torch.cuda.synchronize()
    sampler.start()
    eng.timer_reset()
    wall = time.time()
    step_wall = []
    for _ in range(args.steps):
        t_step = time.perf_counter()
        out = run_query(lib, q)  # ends with the checksum read-back, i.e. synchronised
        step_wall.append(round(1e3 * (time.perf_counter() - t_step), 3))
    ms, launches = eng.timer_read()
    wall = time.time() - wall
    assert out == want
    ms_per_step = ms / args.steps
    value = 2 * rows / (ms_per_step / 1e3)
    # the same K steps again with CUDA events around every kernel launch (adds
    # ~1.5 % to a step, so it is kept out of `value`): per-kernel times for the roofline
    eng.profile(True)
    eng.timer_reset()
    for _ in range(args.steps):
        out = run_query(lib, q)
    ms_prof, _ = eng.timer_read()
    prof_all = eng.profile_read()
    prof = {k: v for k, v in prof_all.items() if not k.startswith("gap_before")}
    eng.profile(False)
    assert out == want

    # ---- e2e: host buffers in, checksums out, every step
    for _ in range(min(args.warmup, 2)):
        upload_all(); run_query(lib, q)
    eng.sync()
    eng.timer_reset()
    for _ in range(args.steps):
        upload_all()
        out = run_query(lib, q)
    ms_e2e, _ = eng.timer_read()
    clocks = sampler.stop()
    assert out == want
    e2e_value = 2 * rows / (ms_e2e / args.steps / 1e3)

    # ---- roofline of the dominant kernel: the one-sweep radix pass on packed
    # 8-byte tuples: 8 B read + 8 B written per tuple per launch;
    # tuples per step = 4 passes x (filtered lhs run + 100M-row rhs run)
    # sizes of the intermediate runs (one untimed pass over the primitives)
    ids = eng.filter_scan(0, 2, ">", 500000)
    lhs = eng.rowids_count(ids)
    tl, tr = eng.build_tuples(0, 1, ids), eng.build_tuples(1, 1)
    eng.sort_tuples(tl); eng.sort_tuples(tr)
    o_l, o_r = eng.merge_join(tl, tr)
    pairs = eng.rowids_count(o_l)
    for h in (ids, o_l, o_r):
        eng.rowids_free(h)
    eng.tuples_free(tl); eng.tuples_free(tr)
    _, key_max = eng.column_info(1, 1)
    passes = (max(1, int(key_max).bit_length()) + 7) // 8
    sort_tuples = lhs + rows
    # algorithmic HBM bytes per STEP of every kernel that can dominate:
    # packed 8-byte tuples, 4-byte row ids, 8-byte column values
    models = {
        "msd_partition": (16.0 * sort_tuples * 2, "16 B per tuple per launch (8 read + 8 written), two partition levels"),
        "msd_count_sort": (16.0 * sort_tuples, "16 B per tuple (8 read + 8 written)"),
        "msd_hist": (8.0 * sort_tuples * 2, "8 B per tuple per level"),
        "onesweep_k": (16.0 * sort_tuples * passes, "16 B per tuple per pass (8 read + 8 written; the reference's 16-byte tuples would be 32 B)"),
        "checksum": (12.0 * pairs * 3, "4 B row id + 8 B value per row per projected column"),
        "join_bounds": (8.0 * sort_tuples + 8.0 * lhs, "8 B per input tuple + 8 B (lb,cnt) per lhs tuple"),
        "join_write": (8.0 * lhs + 8.0 * pairs + 4.0 * pairs, "8 B (lb,cnt) per lhs tuple + 8 B per pair written + 4 B rhs id per pair"),
        "build_tuples": (8.0 * rows + 12.0 * lhs + 8.0 * sort_tuples, "8 B key (+4 B id) in, 8 B packed tuple out"),
    }
    total_kernel_ms = sum(v["ms"] for v in prof.values())
    dom = max((k for k in prof if k in models), key=lambda k: prof[k]["ms"])
    dprof = prof[dom]
    alg_bytes = models[dom][0] * args.steps
    achieved = alg_bytes / (dprof["ms"] / 1e3) / 1e9 if dprof["ms"] else 0.0
    traffic = None  # DRAM bytes per launch from the committed ncu --set full capture of this kernel
    try:
        tr_db = json.load(open(os.path.join(ROOT, "profiles", "kernel_traffic.json")))
        if dom in tr_db:
            traffic = tr_db[dom]["dram_bytes_per_tuple"] * models[dom][0] / tr_db[dom]["algorithmic_bytes_per_tuple"] \
                / max(1, dprof["launches"] // args.steps)
    except Exception:
        pass
    # whole-query algorithmic bytes in SURVEY 8d's terms (uint64 SoA, 16-byte tuples)
    alg_query = (8 * rows + 8 * lhs) + (32 * lhs + 24 * rows) + (lhs + rows) * (8 + passes * 32) \
        + 16 * (lhs + rows) + 16 * pairs + 3 * 16 * pairs
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": pk["hbm_gbs"], "peak_source": pk_src,
                "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": traffic,
                "algorithmic_bytes_per_launch": alg_bytes / max(1, dprof["launches"]), "launches": dprof["launches"],
                "avg_launch_ms": dprof["ms"] / max(1, dprof["launches"]),
                "share_of_kernel_time": dprof["ms"] / total_kernel_ms if total_kernel_ms else None,
                "algorithmic_bytes": models[dom][1],
                "kernels_ms_per_step": {k: round(v["ms"] / args.steps, 4) for k, v in prof.items()},
                "stream_gaps_ms_per_step": {k[len("gap_before:"):]: round(v["ms"] / args.steps, 4)
                                            for k, v in sorted(prof_all.items(), key=lambda kv: -kv[1]["ms"])
                                            if k.startswith("gap_before") and v["ms"] / args.steps >= 0.01},
                "kernels_frac_of_peak": {k: round(models[k][0] * args.steps / (prof[k]["ms"] / 1e3) / 1e9 / pk["hbm_gbs"], 3)
                                         for k in prof if k in models and prof[k]["ms"]},
                "whole_query": {"algorithmic_gb_survey_8d": alg_query / 1e9, "lhs_rows": lhs, "pairs": pairs,
                                "lsd_passes_for_these_keys": passes,
                                "frac_of_peak": (alg_query / 1e9) / (ms_per_step / 1e3) / pk["hbm_gbs"]}}

    cpu = None
    if not args.no_cpu_baseline:
        try:
            r = time_reference(REF_SAMPLE_ROWS, 2, 0)
            if r:
                cpu = {"value": r["rows_per_s"], "unit": "rows/s", "cores": 1, "kind": "reference",
                       "sample": "oracle/_ref/queries (unmodified reference, gcc -O2) on the C2 scaled twin: 2 x %d rows, "
                                 "mean of 2 runs, process wall minus load-only run; host has %d cores, the reference "
                                 "uses 1" % (REF_SAMPLE_ROWS, nproc)}
        except Exception as e:  # the baseline is reported, never required
            cpu = {"value": None, "unit": "rows/s", "cores": 1, "kind": "reference", "sample": "failed: %r" % (e,)}

    line = {
        "metric": METRIC, "value": value, "unit": "rows/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic (numpy default_rng seeds 1/2, generated in %.1f s)" % gen_s,
        "config": {"workload": "C2: single 2-way equi-join + range filter, 2 x %d-row uint64 relations x 3 columns, "
                               "query %s" % (rows, q.strip()), "rows_per_relation": rows,
                   "l2": "inputs larger than L2: %.1f GB of columns and >= 1.2 GB of tuples touched per step" % (6 * col_bytes / 1e9),
                   "result": want.strip(), "queries_per_s": 1e3 / ms_per_step},
        "e2e": {"value": e2e_value, "unit": "rows/s", "h2d_bytes_per_step": 6 * col_bytes, "d2h_bytes_per_step": 3 * 8 + 5 * 16,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "wall_ms_per_step": 1e3 * wall / args.steps, "step_wall_ms": step_wall, "ms_per_step_with_kernel_events": ms_prof / args.steps,
    }
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())

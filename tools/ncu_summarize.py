"""Summarise an .ncu-rep (ncu --set full) into JSON: per launch the duration, DRAM bytes,
throughput percentages, occupancy, pipe utilisation and stall reasons per issue.
Usage: python tools/ncu_summarize.py gpurun_out/x.ncu-rep [more.ncu-rep ...] > profiles/x.json"""
import csv, io, json, subprocess, sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "nvlrx__bytes.sum", "nvltx__bytes.sum"]
STALL = "smsp__average_warps_issue_stalled_"


def summarize(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(head)}
    out = []
    for r in body:
        k = {"Kernel Name": r[idx["Kernel Name"]]}
        for m in KEEP:
            if m in idx:
                k[m] = (r[idx[m]] + " " + units[idx[m]]).strip()
        stalls = {}
        for h, i in idx.items():
            if h.startswith(STALL) and h.endswith("_per_issue_active.ratio"):
                try:
                    stalls[h[len(STALL):-len("_per_issue_active.ratio")]] = round(float(r[i]), 3)
                except ValueError:
                    pass
        k["stalls_per_issue_top"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
        out.append(k)
    return out


if __name__ == "__main__":
    print(json.dumps({p: summarize(p) for p in sys.argv[1:]}, indent=1))

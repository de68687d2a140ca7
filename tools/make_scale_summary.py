"""profiles/scale_rN.json from the per-N bench lines (profiles/bench_rN_n{1,2,4,8}.json): the four configs side by side.
Usage: python tools/make_scale_summary.py r2 > profiles/scale_r2.json"""
import json, os, sys
tag = sys.argv[1]
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
out = {"note": "device time, max over ranks; C2/C3/C4 weak scaling (rows per relation per GPU fixed), C5 strong (same 1000-query batch)", "n": {}}
base = {}
for n in (1, 2, 4, 8):
    p = os.path.join(root, "bench_%s_n%d.json" % (tag, n))
    if not os.path.exists(p):
        continue
    d = json.loads(open(p).read().strip().splitlines()[-1])
    row = {"c2": {"ms_per_step": round(d["ms_per_step"], 3), "G_rows_per_s": round(d["value"] / 1e9, 2), "e2e_ms_per_step": round(d["e2e"]["ms_per_step"], 1),
                  "parity": d["parity"]["twin_vs_reference"] and d["parity"]["full_vs_checker"],
                  "kernels_ms": {k: v for k, v in list(d["roofline"]["kernels_ms_per_step"].items())[:8]}}}
    for c, v in d.get("configs", {}).items():
        if "error" in v and v["error"]:
            row[c] = {"error": v["error"]}
            continue
        row[c] = {"ms_per_step": round(v["ms_per_step"], 2), "G_rows_per_s": round(v["value"] / 1e9, 2), "queries_per_s": round(v["queries_per_s"], 1),
                  "placement": v.get("placement"), "parity": v["parity"]["twin_vs_reference"] and v["parity"]["full_vs_checker"]}
    if n == 1:
        base = row
    else:
        row["c2"]["vs_n_x_single_gpu"] = round(row["c2"]["G_rows_per_s"] / (n * base["c2"]["G_rows_per_s"]), 3)
        for c in ("c3", "c4"):
            if c in row and c in base and "G_rows_per_s" in row[c]:
                row[c]["vs_n_x_single_gpu"] = round(row[c]["G_rows_per_s"] / (n * base[c]["G_rows_per_s"]), 3)
        if "c5" in row and "c5" in base and "queries_per_s" in row["c5"]:
            row["c5"]["speedup_vs_1_gpu"] = round(row["c5"]["queries_per_s"] / base["c5"]["queries_per_s"], 2)
    out["n"][str(n)] = row
print(json.dumps(out, indent=1))

"""ncu launch list (--csv --log-file with gpu__time_duration.sum [+ dram / tensor / lts metrics]) -> per-kernel summary.

    python scripts/summarize_launches.py launches.csv [--skip N] [--traffic out.json]

Prints, per kernel name: launches, total / mean duration and share of the listed launches, and (when the metrics were
collected) DRAM bytes and the time-weighted tensor-pipe activity.  --traffic also maps the launches of the LAST encoder
forward in the list to layers (a layer ends with its k_knn_edge launch) and writes the DRAM bytes of the fused
kNN+EdgeConv path per layer: the file bench.py reads for roofline.traffic."""
import collections
import csv
import json
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = collections.OrderedDict()
    for r in csv.DictReader(lines):
        d = rows.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)  # -> us
        if u in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        d[r["Metric Name"]] = v
    return list(rows.values())


def short(name):
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    return re.sub(r"\(.*", "", name)


def main():
    path = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 0
    rows = load(path)[skip:]
    T, TP, DR, DW = "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum"
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(short(r["name"]), {"n": 0, "us": 0.0, "dram": 0.0, "tp": 0.0})
        a["n"] += 1
        a["us"] += r.get(T, 0.0)
        a["dram"] += r.get(DR, 0.0) + r.get(DW, 0.0)
        a["tp"] += r.get(TP, 0.0) * r.get(T, 0.0)
    tot = sum(a["us"] for a in agg.values())
    print(f"{len(rows)} launches, {tot / 1e3:.3f} ms of serialised kernel time")
    print(f"{'kernel':44s} {'n':>4s} {'total us':>10s} {'mean us':>9s} {'share':>7s} {'DRAM MB':>10s} {'tensor %':>9s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        print(f"{k:44s} {a['n']:4d} {a['us']:10.1f} {a['us'] / a['n']:9.1f} {100 * a['us'] / tot:6.1f}% {a['dram'] / 1e6:10.1f} "
              f"{(a['tp'] / a['us'] if a['us'] else 0):9.1f}")
    if "--traffic" in sys.argv:
        out = sys.argv[sys.argv.index("--traffic") + 1]
        # last forward = everything after the last k_normalize launch
        starts = [i for i, r in enumerate(rows) if "k_normalize" in r["name"]]
        fwd = rows[starts[-1]:]
        layer, acc, res = 0, {"dram": 0.0, "us": 0.0}, {}
        for r in fwd:
            n = short(r["name"])
            if any(k in n for k in ("k_knn_pack", "k_knn_tc", "k_knn_rerank", "k_knn_small", "k_knn_edge")):
                acc["dram"] += r.get(DR, 0.0) + r.get(DW, 0.0)
                acc["us"] += r.get(T, 0.0)
            if "k_knn_edge" in n:
                res[f"layer{layer}_dram_bytes_per_launch"] = int(acc["dram"])
                res[f"layer{layer}_ncu_us"] = round(acc["us"], 1)
                layer, acc = layer + 1, {"dram": 0.0, "us": 0.0}
        res = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per layer of the fused kNN+EdgeConv path (k_knn_pack + k_knn_tc + "
                           "k_knn_rerank + k_knn_edge; k_knn_small_tiled + k_knn_edge for layers 5-6), one ncu pass over scripts/one_forward.py "
                           "(B=256, N=1024, shipped weights), serialised cold-cache launches; read by bench.py for roofline.traffic. "
                           f"Source: {path} via scripts/summarize_launches.py", **res}
        json.dump(res, open(out, "w"), indent=1)
        print("wrote", out)


if __name__ == "__main__":
    main()

"""gpurun_out/r2final/*.raw.csv / *.source.csv (written by scripts/gpu_round2_final.sh) -> profiles/r02/ncu_full_key_metrics.csv
and profiles/r02/ncu_source_hotspots.txt."""
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O = os.path.join(ROOT, "gpurun_out", "r2final") + "/"
P = os.path.join(ROOT, "profiles", "r02") + "/"
COLS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def I(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


def short(n):
    return re.sub(r"\(.*", "", n.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", ""))


def key_metrics():
    out = open(P + "ncu_full_key_metrics.csv", "w", newline="")
    w = csv.writer(out)
    first = True
    for f, tag in (("prof_gemm.raw.csv", "one_forward.py (2nd forward), GEMMs"), ("prof_knn.raw.csv", "one_forward.py (2nd forward), kNN + EdgeConv"),
                   ("prof_sdf.raw.csv", "one_sdf.py, SDF GEMMs (131072 columns)")):
        rows = list(csv.reader(open(O + f)))
        hdr, units = rows[0], rows[1]
        use = [c for c in COLS if c in hdr]
        if first:
            w.writerow(["capture", "kernel", "grid", "block"] + [c + " [" + units[hdr.index(c)] + "]" for c in use])
            first = False
        for r in rows[2:]:
            w.writerow([tag, short(r[hdr.index("Kernel Name")]), r[hdr.index("Grid Size")], r[hdr.index("Block Size")]] + [r[hdr.index(c)] for c in use])
            print(short(r[hdr.index("Kernel Name")])[:30].ljust(30), r[hdr.index("Grid Size")].ljust(13),
                  "us", r[hdr.index(COLS[0])][:7], "tensor%", r[hdr.index(COLS[1])][:5], "issue%", r[hdr.index(COLS[3])][:5])
    out.close()


def hotspots():
    out = ["ncu --set full --import-source on, source page (SASS view), captured at the round-2 final commit by scripts/gpu_round2_final.sh", ""]
    for f in ("prof_knn_tc_layer1.source.csv", "prof_knn_edge_layer2.source.csv"):
        rows = list(csv.reader(open(O + f)))
        kname = rows[0][1]
        hdr = rows[1]
        i_src, i_s, i_ex, i_a = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Address")
        seen, body = set(), []
        for r in rows[2:]:
            if len(r) <= i_ex or r[i_a] in seen:
                continue
            seen.add(r[i_a])
            body.append(r)
        stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {h: sum(I(r[hdr.index(h)]) for r in body if hdr.index(h) < len(r)) for h in stall}
        tot = sum(I(r[i_s]) for r in body)
        out.append(f"== {kname}: {tot} warp-state samples over {len(body)} SASS instructions")
        out.append("stall reasons: " + ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
        out.append("hottest instructions (samples, executions, SASS, top stall):")
        for k in sorted(range(len(body)), key=lambda k: -I(body[k][i_s]))[:10]:
            r = body[k]
            st = max(stall, key=lambda h: I(r[hdr.index(h)]) if hdr.index(h) < len(r) else 0)
            out.append(f"  {I(r[i_s]):6d} {I(r[i_ex]):9d}  {r[i_src].strip()[:72]:72s} {st[6:]}")
        out.append("")
    open(P + "ncu_source_hotspots.txt", "w").write("\n".join(out))


if __name__ == "__main__":
    key_metrics()
    hotspots()
    print("wrote", P + "ncu_full_key_metrics.csv", "and", P + "ncu_source_hotspots.txt")

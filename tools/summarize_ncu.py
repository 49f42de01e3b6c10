"""Condense `ncu --page raw --csv` exports into the per-kernel summary tables kept under profiles/.
usage: python tools/summarize_ncu.py <raw.csv> [<raw.csv> ...] --out profiles/ncu_full_<tag>.csv [--traffic profiles/traffic_<tag>.json]"""
import argparse
import csv
import json

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw", nargs="+")
    ap.add_argument("--out", required=True)
    ap.add_argument("--traffic")
    a = ap.parse_args()
    out_rows, units, traffic = [], None, {}
    for f in a.raw:
        rows = list(csv.reader(open(f)))
        h = rows[0]
        idx = {c: h.index(c) for c in COLS if c in h}
        units = ["", "", "", ""] + [rows[1][idx[c]] if c in idx else "" for c in COLS]
        for r in rows[2:]:
            name = r[h.index("Kernel Name")]
            out_rows.append([r[h.index("ID")], name, r[h.index("Block Size")], r[h.index("Grid Size")]]
                            + [r[idx[c]] if c in idx else "" for c in COLS])
            rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            ur, uw = rows[1][idx["dram__bytes_read.sum"]], rows[1][idx["dram__bytes_write.sum"]]
            short = name.replace("void ", "").split("(")[0]
            traffic.setdefault(short, []).append(rd * scale.get(ur, 1.0) + wr * scale.get(uw, 1.0))
    with open(a.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["ID", "Kernel Name", "Block Size", "Grid Size"] + COLS)
        w.writerow(units)
        w.writerows(out_rows)
    if a.traffic:
        json.dump({k: sum(v) / len(v) for k, v in traffic.items()}, open(a.traffic, "w"), indent=1)
    print("wrote", a.out, len(out_rows), "kernels")


if __name__ == "__main__":
    main()

"""Condense `ncu --page raw --csv` exports into the per-launch figures quoted in DESIGN.md / profiles/README.md.
  python scripts/ncu_summary.py profiles/ncu_kblock_r2_raw.csv [--traffic profiles/kblock_traffic_r2.json]
Prints one JSON object per profiled launch; with --traffic also writes the DRAM bytes per launch of the
k_block launches (block 0, 1, ... in launch order) in the form bench.py reads for `roofline.traffic`."""
import csv
import json
import sys

KEEP = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "shared_wavefronts_pct_of_peak",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "launch__registers_per_thread": "registers",
    "launch__stack_size": "stack_bytes",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}


def main():
    path = sys.argv[1]
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0]}
        for name, key in KEEP.items():
            if name not in hdr:
                continue
            i = hdr.index(name)
            try:
                v = float(r[i])
            except ValueError:
                continue
            if v != v:
                continue
            u = units[i]
            if key == "time":
                d["time_ms"] = v * SCALE.get(u, 1.0) * 1e3
            elif key in ("dram_read", "dram_write"):
                d[key + "_bytes"] = v * SCALE.get(u, 1.0)
            else:
                d[key] = v
        if "dram_read_bytes" in d and "dram_write_bytes" in d and d.get("time_ms"):
            d["dram_bytes"] = d["dram_read_bytes"] + d["dram_write_bytes"]
            d["dram_gbs"] = d["dram_bytes"] / (d["time_ms"] * 1e-3) / 1e9
        out.append(d)
        print(json.dumps(d))
    if "--traffic" in sys.argv:
        tpath = sys.argv[sys.argv.index("--traffic") + 1]
        kb = [d for d in out if d["kernel"].endswith("k_block") or "k_block" in d["kernel"]]
        json.dump({f"k_block(block={i})": d["dram_bytes"] for i, d in enumerate(kb)} |
                  {"source": path, "note": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, bench workload h1m (1,000,000 rows, K=20), steady-state sweep"},
                  open(tpath, "w"), indent=1)


if __name__ == "__main__":
    main()

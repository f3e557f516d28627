"""Writes profiles/r02_ncu_traffic.json (what bench.py's roofline.traffic is derived from) from an `ncu --set full` capture:

    ncu -i gpurun_out/<capture>.ncu-rep --page raw --csv > profiles/<capture>_raw.csv
    python tools/ncu_traffic.py profiles/<capture>_raw.csv <proposals evaluated in the profiled launch> "<command that was profiled>"
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    raw, proposals, command = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    rows = list(csv.reader(open(raw)))
    head, units, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(head)}

    def get(name, want_unit):
        v, u = float(vals[col[name]].replace(",", "")), units[col[name]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "%": 1.0, "": 1.0}
        return v * scale.get(u, 1.0)

    out = {"source": os.path.relpath(raw, ROOT), "command": command, "kernel": vals[col["Kernel Name"]],
           "duration_ms": get("gpu__time_duration.sum", "ms"), "dram_bytes_read": get("dram__bytes_read.sum", "byte"),
           "dram_bytes_write": get("dram__bytes_write.sum", "byte"), "proposals_in_launch": proposals,
           "l2_hit_pct": get("lts__t_sector_hit_rate.pct", "%"), "l1_hit_pct": get("l1tex__t_sector_hit_rate.pct", "%"),
           "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active", "%"),
           "threads_per_instruction": get("smsp__thread_inst_executed_per_inst_executed.ratio", ""),
           "warp_instructions": get("smsp__inst_executed.sum", ""), "registers_per_thread": get("launch__registers_per_thread", ""),
           "grid": get("launch__grid_size", "")}
    out["dram_bytes_per_proposal"] = (out["dram_bytes_read"] + out["dram_bytes_write"]) / proposals
    out["warp_instructions_per_proposal"] = out["warp_instructions"] / proposals
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

"""Summarise an `ncu --set full` capture (exported with `ncu -i X.ncu-rep --page raw --csv > raw.csv`) into
profiles/<tag>_ncu_summary.{json,md}: one row per profiled launch with the metrics the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/<run>/raw.csv <tag> ["free-text note"]
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("duration_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("regs", "launch__registers_per_thread"),
    ("smem_dyn_B", "launch__shared_mem_per_block_dynamic"),
    ("blocks_per_sm_limit_smem", "launch__occupancy_limit_shared_mem"),
    ("blocks_per_sm_limit_regs", "launch__occupancy_limit_registers"),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("warp_inst", "smsp__inst_executed.sum"),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    ("threads_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"),
    ("sm_cycles_active_avg", "sm__cycles_active.avg"),
    ("sm_cycles_active_max", "sm__cycles_active.max"),
    ("fma_pipe_pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
    ("fp64_pipe_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
    ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("dram_read_B", "dram__bytes_read.sum"),
    ("dram_write_B", "dram__bytes_write.sum"),
    ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("lsu_wavefronts_pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    ("stall_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}


def main():
    raw, tag = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for d in data:
        rec = {"kernel": d[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for name, key in KEYS:
            if key not in hdr:
                rec[name] = None
                continue
            i = hdr.index(key)
            try:
                v = float(d[i].replace(",", ""))
            except ValueError:
                rec[name] = None
                continue
            v *= UNIT.get(units[i], 1.0) if name.endswith("_B") or name == "duration_us" else 1.0
            rec[name] = round(v, 3)
        rec["dram_traffic_B"] = (rec["dram_read_B"] or 0) + (rec["dram_write_B"] or 0)
        out.append(rec)
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    base = os.path.join(ROOT, "profiles", tag + "_ncu_summary")
    json.dump({"source": raw, "note": note, "launches": out}, open(base + ".json", "w"), indent=1)
    with open(base + ".md", "w") as f:
        f.write("# ncu --set full summary: %s\n\n%s\n\nSource capture: `%s` (cold-cache, serialised replays: compare shares, not absolutes).\n\n" % (tag, note, raw))
        names = ["kernel"] + [k for k, _ in KEYS] + ["dram_traffic_B"]
        f.write("| metric | " + " | ".join("%d: %s" % (i, r["kernel"].replace("rtk::", "")) for i, r in enumerate(out)) + " |\n")
        f.write("|---|" + "---|" * len(out) + "\n")
        for n in names[1:]:
            f.write("| %s | " % n + " | ".join(str(r[n]) for r in out) + " |\n")
    print("wrote", base + ".json/.md")


if __name__ == "__main__":
    main()

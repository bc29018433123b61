#!/usr/bin/env python3
"""Print the metrics we track from an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv, subprocess, sys, io
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread ", "launch__occupancy_limit_registers",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "per_issue_active.ratio",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum ", "dram__bytes_write.sum ",
        "smsp__inst_executed.sum ", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg ", "sass__inst_executed_global_loads",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
def summary(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    out = []
    for r in rows[2:]:
        out.append("### " + r[rows[0].index("Kernel Name")][:60])
        for h, u, v in zip(rows[0], rows[1], r):
            if any(k in h + " " for k in KEYS):
                try:
                    if "per_issue" in h and float(v) < 0.03:
                        continue
                except ValueError:
                    pass
                out.append("| %s | %s | %s |" % (h, u, v))
    return "\n".join(out)
def traffic(path):
    """kernel name -> dram bytes (read + write) of the first captured launch of each kernel"""
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, u = rows[0], rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    out = {}
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].split("<")[0].replace("void ", "").replace("b381::", "")
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = h.index(key)
            tot += float(r[i]) * scale.get(u[i], 1)
        out.setdefault(name, tot)
    return out
if __name__ == "__main__":
    if sys.argv[1] == "--traffic":
        import json
        t = {}
        for p in sys.argv[2:]:
            t.update(traffic(p))
        print(json.dumps(t))
        sys.exit(0)
    for p in sys.argv[1:]:
        print("## " + p); print(summary(p))

"""ncu raw page (csv) with SEVERAL kernel launches -> one json list (duration, DRAM traffic, instructions, occupancy, issue rate)
   ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary_multi.py raw.csv "<what was run>" > profiles/NAME.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
def val(r, n, scale=True):
    if n not in h: return None
    i = h.index(n)
    try: x = float(r[i].replace(",", ""))
    except ValueError: return r[i]
    if scale: x *= {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "nsecond": 1e-3}.get(units[i], 1.0)
    return x
out = []
for r in rows[2:]:
    if len(r) < len(h): continue
    out.append({"kernel": r[h.index("Kernel Name")], "grid": r[h.index("Grid Size")], "block": r[h.index("Block Size")],
                "duration_us": val(r, "gpu__time_duration.sum"),
                "dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
                "l2_sector_hit_rate_pct": val(r, "lts__t_sector_hit_rate.pct", False),
                "registers_per_thread": val(r, "launch__registers_per_thread", False),
                "warp_instructions": val(r, "smsp__inst_executed.sum", False),
                "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False),
                "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False),
                "sm_throughput_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed", False),
                "shared_bank_conflicts": val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", False),
                "shared_wavefronts": val(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False)})
print(json.dumps({"source": sys.argv[2], "launches": out}, indent=1))

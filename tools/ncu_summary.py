"""ncu raw page (csv) of one kernel launch -> the small json bench.py quotes as roofline.traffic
   ncu -i X.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv q4_0 514 "<command>" > profiles/NAME.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
h, v = rows[0], rows[2]
g = {n: v[i] for i, n in enumerate(h)}
u = {n: rows[1][i] for i, n in enumerate(h)}
def val(n, scale_units=True):
    x = float(g[n].replace(",", ""))
    if scale_units:
        x *= {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u[n], 1.0)
    return x
out = {"kernel": g["Kernel Name"], "ftype": sys.argv[2], "n_past": int(sys.argv[3]), "source": sys.argv[4],
       "duration_us": val("gpu__time_duration.sum"),
       "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
       "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
       "l2_sector_hit_rate_pct": val("lts__t_sector_hit_rate.pct", False),
       "registers_per_thread": val("launch__registers_per_thread", False),
       "sm_throughput_pct": val("sm__throughput.avg.pct_of_peak_sustained_elapsed", False),
       "lts_throughput_pct": val("lts__throughput.avg.pct_of_peak_sustained_elapsed", False),
       "icc_hit_rate_pct": val("sm__icc_request_hit_rate.pct", False)}
print(json.dumps(out, indent=1))

"""Condense `ncu -i X.ncu-rep --page raw --csv` into per-kernel averages (JSON on stdout).

usage: ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
       python tools/ncu_summary.py raw.csv <workload> <microbes_per_gpu> > profiles/ncu_summary.json
"""
import collections
import csv
import json
import sys

WANT = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "registers",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_instruction",
}
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6, "msecond": 1e3,
              "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
acc = collections.OrderedDict()
for r in rows[2:]:
    name = r[ik].replace("<unnamed>::", "").split("(")[0].replace("void ", "").split("<")[0].replace("lm::", "")
    a = acc.setdefault(name, collections.defaultdict(list))
    for m, key in WANT.items():
        if m in hdr:
            i = hdr.index(m)
            try:
                a[key].append(float(r[i].replace(",", "")) * UNIT_SCALE.get(units[i], 1.0))
            except ValueError:
                pass
out = {}
for name, a in acc.items():
    out[name] = {k: sum(v) / len(v) for k, v in a.items()}
    out[name]["launches_captured"] = len(a["time_us"])
    out[name]["microbes_per_gpu"] = int(sys.argv[3])
print(json.dumps({sys.argv[2]: out}, indent=1))

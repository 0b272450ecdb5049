"""`ncu -i X.ncu-rep --page raw --csv` -> per-kernel DRAM / L2 traffic per launch (JSON on stdout), the source of
`roofline.traffic` in bench.py.    ncu -i gpurun_out/r2_full.ncu-rep --page raw --csv > /tmp/raw.csv; python profiles/traffic.py /tmp/raw.csv"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
ki = hdr.index("Kernel Name")


def col(name):
    return hdr.index(name) if name in hdr else None


c_rd, c_wr, c_t, c_l2 = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum"), col("lts__t_bytes.sum")
units = rows[1]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_us(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(u, 1)


out = {}
for r in rows[2:]:
    m = re.search(r"(\w+)_kernel", r[ki])
    name = m.group(1) if m else r[ki].split("::")[-1].split("(")[0].split("<")[0]
    e = out.setdefault(name, {"launches_captured": 0, "dram": 0.0, "l2": 0.0, "us": 0.0})
    e["launches_captured"] += 1
    e["dram"] += to_bytes(r[c_rd], units[c_rd]) + to_bytes(r[c_wr], units[c_wr])
    if c_l2 is not None:
        e["l2"] += to_bytes(r[c_l2], units[c_l2])
    e["us"] += to_us(r[c_t], units[c_t])
res = {"source": "ncu --set full --clock-control none over the launches of one training step (eager launches, B = 64); "
                 "dram__bytes_read.sum + dram__bytes_write.sum (and lts__t_bytes.sum) per launch",
       "kernels": {k: {"launches_captured": v["launches_captured"], "dram_bytes_per_launch": int(v["dram"] / v["launches_captured"]),
                       "l2_bytes_per_launch": int(v["l2"] / v["launches_captured"]), "ncu_us_total": round(v["us"], 2)}
                   for k, v in out.items()}}
print(json.dumps(res, indent=1))

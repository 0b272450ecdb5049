"""Turn `ncu -i X.ncu-rep --page raw --csv` output into the compact per-kernel table committed under profiles/.

    ncu -i gpurun_out/r1a_prof.ncu-rep --page raw --csv > /tmp/raw.csv && python profiles/summarize.py /tmp/raw.csv
"""
import csv
import re
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rd_MB"), ("dram__bytes_write.sum", "wr_MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("lts__t_bytes.sum", "l2_bytes")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    idx = [(hdr.index(c), n, units[hdr.index(c)]) for c, n in COLS if c in hdr]
    print("| # | kernel | " + " | ".join(f"{n} ({u})" if u and n not in ("us",) else n for _, n, u in idx) + " |")
    print("|---|---|" + "---|" * len(idx))
    for i, r in enumerate(rows[2:]):
        m = re.search(r"(\w+_kernel)(<[^>]*>)?", r[ki])
        name = (m.group(1) + (m.group(2) or "")) if m else r[ki].split("::")[-1].split("(")[0]
        name = name.replace("(int)", "").replace("(bool)", "").replace("__nv_bfloat16", "bf16")[:44]
        vals = []
        for j, n, u in idx:
            try:
                vals.append(f"{float(r[j].replace(',', '')):.2f}")
            except ValueError:
                vals.append(r[j])
        print(f"| {i} | {name} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])

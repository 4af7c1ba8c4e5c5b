"""Summarises ncu --set full reports (one kernel launch each) into one CSV row per report:
duration, DRAM bytes, tensor-pipe activity, L2 hit rate, occupancy, registers.
Usage: python tools/summarize_ncu.py out.csv report1.ncu-rep [report2.ncu-rep ...]"""
import csv
import io
import os
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_operand_mem_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_MB"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__cycles_elapsed.max", "sm_cycles"),
]


def to_base(value, unit):
    v = float(value.replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
    return v * scale.get(unit, 1.0)


def main():
    out_path, reports = sys.argv[1], sys.argv[2:]
    rows = []
    for rep in reports:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        table = list(csv.reader(io.StringIO(raw)))
        if len(table) < 3:
            continue
        hdr, units = table[0], table[1]
        for r in table[2:]:
            d = {"report": os.path.basename(rep), "kernel": r[hdr.index("Kernel Name")].split("(")[0]}
            for m, name in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    try:
                        d[name] = round(to_base(r[i], units[i]), 3)
                    except ValueError:
                        d[name] = r[i]
            rows.append(d)
    cols = ["report", "kernel"] + [n for _, n in METRICS]
    with open(out_path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for d in rows:
            w.writerow(d)
    for d in rows:
        print(d)


if __name__ == "__main__":
    main()

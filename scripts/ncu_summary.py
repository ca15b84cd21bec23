"""Condenses `ncu --set full` reports (.ncu-rep) into a small CSV + markdown table for profiles/.

usage: python scripts/ncu_summary.py OUT_PREFIX report1.ncu-rep [report2.ncu-rep ...]
Needs the `ncu` CLI (present in the build container; reads reports without a GPU).
"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration_us"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__waves_per_multiprocessor", "waves"),
]


def read(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"report": report.split("/")[-1], "kernel": r[hdr.index("Kernel Name")].split("(")[0][-60:],
             "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for m, name in METRICS:
            if m in hdr:
                i = hdr.index(m)
                d[name] = f"{r[i]} {units[i]}".strip()
        res.append(d)
    return res


def main():
    prefix, reports = sys.argv[1], sys.argv[2:]
    rows = [d for rep in reports for d in read(rep)]
    cols = ["report", "kernel", "grid", "block"] + [n for _, n in METRICS]
    with open(prefix + ".csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=cols)
        w.writeheader()
        for d in rows:
            w.writerow({c: d.get(c, "") for c in cols})
    with open(prefix + ".md", "w") as f:
        f.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
        for d in rows:
            f.write("| " + " | ".join(str(d.get(c, "")) for c in cols) + " |\n")
    print(open(prefix + ".md").read())


if __name__ == "__main__":
    main()

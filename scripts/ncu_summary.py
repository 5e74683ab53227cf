"""Summarise an .ncu-rep (read on the CPU box): python scripts/ncu_summary.py rep [--md out.md] [--grep regex]"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max",
    "sm__cycles_active.avg",
    "launch__grid_size",
    "launch__registers_per_thread",
    "sm__inst_executed_pipe_tensor.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_l1tex2xbar_write_bytes.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg",
]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    md = sys.argv[sys.argv.index("--md") + 1] if "--md" in sys.argv else None
    pat = re.compile(sys.argv[sys.argv.index("--grep") + 1]) if "--grep" in sys.argv else None
    hdr, units, data = load(rep)
    col = {h: i for i, h in enumerate(hdr)}
    names = [d[col["Kernel Name"]][:48] for d in data]
    lines = ["| metric | unit | " + " | ".join(f"#{i} {n}" for i, n in enumerate(names)) + " |",
             "|---|---|" + "---|" * len(names)]
    keys = [m for m in METRICS if m in col]
    if pat:
        keys += [h for h in hdr if pat.search(h) and h not in keys]
    for m in keys:
        i = col[m]
        lines.append(f"| {m} | {units[i]} | " + " | ".join(d[i] for d in data) + " |")
    text = "\n".join(lines)
    print(text)
    if md:
        open(md, "w").write(text + "\n")


if __name__ == "__main__":
    main()

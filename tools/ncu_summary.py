"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown table for profiles/."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    cols = [k for k in KEYS if k in hdr]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {rep}\n\n| kernel | " + " | ".join(cols) + " |\n|" + "---|" * (len(cols) + 1) + "\n")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0]
            f.write("| " + name + " | " + " | ".join(f"{r[hdr.index(c)]} {units[hdr.index(c)]}" for c in cols) + " |\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

#!/usr/bin/env python
"""Extract the metrics the roofline discussion needs from an .ncu-rep into markdown + JSON under profiles/.
Usage: ncu_summary.py <report.ncu-rep> <out_basename> [title]"""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__cluster_size", "smsp__inst_executed.sum"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:90]}
        for i, h in enumerate(hdr):
            if h in WANT:
                d[h] = f"{r[i]} {units[i]}".strip()
        launches.append(d)
    json.dump({"title": title, "launches": launches}, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none)\n\n")
        for i, d in enumerate(launches):
            f.write(f"## launch {i}: `{d['kernel']}`\n\n| metric | value |\n|---|---|\n")
            for k in WANT:
                if k in d:
                    f.write(f"| {k} | {d[k]} |\n")
            f.write("\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main()

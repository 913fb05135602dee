#!/usr/bin/env python
"""Turn `ncu -i X.ncu-rep --page raw --csv` output into a small markdown table of the metrics the roofline argument needs
(duration, DRAM bytes / throughput %, L2 throughput %, tensor-pipe active %, achieved occupancy, registers, shared memory)."""
import csv, sys

KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %peak"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %peak"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def main():
    src, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
    rows = [r for r in csv.reader(open(src)) if r]
    hdr = next(r for r in rows if "Kernel Name" in r)
    hi = rows.index(hdr)
    units = rows[hi + 1]
    ix = {h: i for i, h in enumerate(hdr)}
    print("# %s\n" % title)
    print("| kernel | " + " | ".join(n for _, n in KEYS) + " |")
    print("|---|" + "---|" * len(KEYS))
    for r in rows[hi + 2:]:
        if len(r) != len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("b200::", "")
        cells = []
        for k, _ in KEYS:
            if k in ix and r[ix[k]] != "":
                cells.append("%s %s" % (r[ix[k]], units[ix[k]]))
            else:
                cells.append("-")
        print("| %s | %s |" % (name[:60], " | ".join(cells)))


if __name__ == "__main__":
    main()

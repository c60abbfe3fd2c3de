#!/usr/bin/env python
"""Per-launch summary table of an `ncu --set full` report (read with `ncu -i ... --page raw --csv`).

    python tools/ncu_summary.py gpurun_out/r02_full_conv_c3.ncu-rep > profiles/r02_ncu_conv_c3.md
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "us", 1.0),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1.0),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %", 1.0),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM thr %", 1.0),
    ("dram__bytes_read.sum", "DRAM read MB", 1.0),
    ("dram__bytes_write.sum", "DRAM write MB", 1.0),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1.0),
    ("launch__registers_per_thread", "regs", 1.0),
    ("launch__shared_mem_per_block_dynamic", "dyn smem KB", 1.0),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"`{path.split('/')[-1]}` (ncu --set full --clock-control none, one B200, captured inside a timed bench step)\n")
    print("| # | kernel | grid | block | " + " | ".join(c[1] for c in COLS) + " |")
    print("|---|---|---|---|" + "---:|" * len(COLS))
    for n, r in enumerate(rows[2:]):
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        if "<" in r[idx["Kernel Name"]]:
            name = r[idx["Kernel Name"]].split("(CUtensorMap")[0].split("(StreamArgs")[0].replace("void ", "").replace("(bool)", "").replace("(int)", "")
        vals = []
        for key, _, _ in COLS:
            v = r[idx[key]] if key in idx else ""
            u = units[idx[key]] if key in idx else ""
            try:
                f = float(v.replace(",", ""))
                if u == "byte":
                    f /= 1e6
                elif u == "Kbyte":
                    f /= 1e3
                elif u == "Gbyte":
                    f *= 1e3
                elif u in ("ns", "nsecond"):
                    f /= 1e3
                elif u in ("ms", "msecond"):
                    f *= 1e3
                if key.startswith("launch__shared") and u == "byte":
                    f *= 1e3
                vals.append(f"{f:.1f}")
            except ValueError:
                vals.append(v)
        print(f"| {n} | `{name[:70]}` | {r[idx['Grid Size']]} | {r[idx['Block Size']]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main(sys.argv[1])

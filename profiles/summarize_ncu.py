#!/usr/bin/env python
"""Turn the ncu artefacts that gpurun brings back into the tracked summaries under profiles/.

  python profiles/summarize_ncu.py gpurun_out/<launches>.csv gpurun_out/<prof>.ncu-rep profiles/<name>.md

The launch list comes from   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python bench.py ...
the full capture from        ncu --set full --clock-control none --import-source on -k regex:... -o ... python bench.py ...
(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes).
"""
import csv
import subprocess
import sys
from collections import defaultdict

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size"]


def launches(path):
    agg = defaultdict(lambda: [0, 0.0])
    for r in csv.reader(open(path)):
        if len(r) > 5 and r[0].isdigit():
            v = float(r[-1].replace(",", ""))
            v = v / 1000 if r[-2] == "ns" else (v * 1000 if r[-2] == "ms" else v)
            name = r[4].split("(")[0]
            agg[name][0] += 1
            agg[name][1] += v
    return agg


def main():
    lcsv, rep, out = sys.argv[1:4]
    lines = ["# ncu summary", "", f"launch list: `{lcsv}`; full capture: `{rep}`", "", "## device time by kernel (all launches of the command)", "",
             "| kernel | launches | total us | share |", "|---|---|---|---|"]
    agg = launches(lcsv)
    tot = sum(v[1] for v in agg.values()) or 1.0
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(raw.splitlines()))
    if rd:
        hdr, units = rd[0], rd[1]
        idx = [(w, hdr.index(w)) for w in WANT if w in hdr]
        kn = hdr.index("Kernel Name")
        lines += ["", "## `--set full` capture (per launch)", "", "| kernel | " + " | ".join(f"{w} [{units[i]}]" for w, i in idx) + " |",
                  "|---|" + "---|" * len(idx)]
        for r in rd[2:]:
            lines.append(f"| `{r[kn].split('(')[0]}` | " + " | ".join(r[i] for _, i in idx) + " |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()

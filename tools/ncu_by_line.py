#!/usr/bin/env python
"""Per-source-line view of an ncu capture without the GUI: joins the SASS page of a kernel in a .ncu-rep
(`ncu --page source --csv`: executed instructions and stall samples per SASS instruction) with the line table of the
same kernel in the object file (`nvdisasm --print-line-info`), by instruction offset.

  python tools/ncu_by_line.py <report.ncu-rep> <object.o> <kernel regex for ncu> [top N] [mangled-name substring for the object]

(template kernels: give the mangled instantiation, e.g. fz_kernelILi1ELi2E, as the last argument)

The object must be the build that was profiled (same SASS); compile with -lineinfo."""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(obj: str, kernel: str):
    d = tempfile.mkdtemp()
    import os
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    import glob
    cub = glob.glob(d + "/*.cubin")[0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
    out, cur, active, inline_stack = {}, None, False, None
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            active = kernel in m.group(1)
            continue
        if ln.startswith("//---") and ".text." not in ln and active and "section" in ln:
            pass
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)), "inlined" in m.group(3))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and cur:
            out[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return out


def main():
    rep, obj, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    lines = sass_lines(obj, sys.argv[5] if len(sys.argv) > 5 else kernel)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    rd = list(csv.reader(raw.splitlines()))
    hi = next(i for i, r in enumerate(rd) if r and r[0] == "Address")
    hdr = rd[hi]
    ci, cs = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    base = None
    per_line = defaultdict(lambda: [0, 0])
    per_op = defaultdict(lambda: [0, 0])
    tot_i = tot_s = 0
    for r in rd[hi + 1:]:
        if len(r) <= cs or not r[0].startswith("0x"):
            if r and r[0] == "Kernel Name":
                break   # next launch of the same kernel
            continue
        a = int(r[0], 16)
        if base is None:
            base = a
        n, s = int(r[ci] or 0), int(r[cs] or 0)
        key = lines.get(a - base, (("?", 0, False), ""))[0]
        per_line[key][0] += n
        per_line[key][1] += s
        op = r[1].split()[0] if r[1].split() else "?"
        if op.startswith("@"):
            op = r[1].split()[1]
        per_op[op.split(".")[0]][0] += n
        per_op[op.split(".")[0]][1] += s
        tot_i += n
        tot_s += s
    print(f"kernel ~{kernel}: {tot_i} warp instructions, {tot_s} stall samples")
    print("-- by source line (instructions %, stall samples %)")
    for k, (n, s) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[0]}:{k[1]:<5} instr {100 * n / max(tot_i, 1):5.1f}%  stalls {100 * s / max(tot_s, 1):5.1f}%")
    print("-- by opcode")
    for k, (n, s) in sorted(per_op.items(), key=lambda kv: -kv[1][0])[:25]:
        print(f"{k:<10} instr {100 * n / max(tot_i, 1):5.1f}%  stalls {100 * s / max(tot_s, 1):5.1f}%")


if __name__ == "__main__":
    main()

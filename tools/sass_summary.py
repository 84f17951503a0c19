"""Static evidence per kernel of libauron_b200.so, generated without a GPU: resource usage (`cuobjdump -res-usage`) and counts of the
SASS instructions that tell how a kernel touches memory (`cuobjdump -sass`): 128-bit global loads / stores, atomics and reductions,
shared-memory traffic, warp shuffles / votes.  Writes profiles/r02_sass_summary.md.

    python tools/sass_summary.py
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "auron_b200", "libauron_b200.so")
PATTERNS = [("LDG.128", r"\bLDG\.E(\.\w+)*\.128"), ("LDG.64", r"\bLDG\.E(\.\w+)*\.64"), ("LDG", r"\bLDG\b"), ("STG.128", r"\bSTG\.E(\.\w+)*\.128"), ("STG", r"\bSTG\b"),
            ("ATOMG/RED", r"\b(ATOMG|RED|ATOM)\b"), ("ATOMS", r"\bATOMS\b"), ("LDS", r"\bLDS\b"), ("STS", r"\bSTS\b"), ("SHFL", r"\bSHFL\b"),
            ("VOTE/MATCH", r"\b(VOTE|MATCH|VOTEU)\b"), ("UBLKCP (TMA bulk copy)", r"\bUBLKCP\b"), ("SYNCS (mbarrier)", r"\bSYNCS\b"), ("BAR", r"\bBAR\b"), ("POPC", r"\bPOPC\b"), ("SHF (funnel)", r"\bSHF\b")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    short = []
    for n in out:
        n = re.sub(r"\(.*", "", n)
        short.append(n.replace("auron::", "").replace("(anonymous namespace)::", ""))
    return short


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(x) for x in m.groups())
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    total = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and re.match(r"\s*/\*[0-9a-f]{4}\*/", line):
            total[cur] += 1
            for name, pat in PATTERNS:
                if re.search(pat, line):
                    counts[cur][name] += 1
    names = sorted(usage)
    pretty = dict(zip(names, demangle(names)))
    rows = []
    for n in names:
        if total[n] < 40:   # trivial helpers
            continue
        r, st, sh, lo = usage[n]
        rows.append((pretty[n], r, st, sh, lo, total[n], counts[n]))
    rows.sort(key=lambda x: x[0])
    with open(os.path.join(ROOT, "profiles", "r02_sass_summary.md"), "w") as f:
        f.write("# Static per-kernel summary (sm_100a SASS of `libauron_b200.so`, `tools/sass_summary.py`)\n\n")
        f.write("Columns: registers per thread, per-thread stack frame in bytes (local arrays: the expression VM's spill-free register file, the scout's "
                "run tables), static shared memory (1024 B are reserved by the system), then counts of SASS instructions by kind.  `LDG.128` / `STG.128` are "
                "the 128-bit vector accesses of the streaming kernels, `ATOMG/RED` the accumulator atomics, `SHFL` / `VOTE` / `POPC` the warp-level rank and "
                "compaction steps, `SHF` the funnel shifts that re-align byte streams.  tcgen05 / TMA instructions do not occur: the path has no dense "
                "contraction and its streams are read once (DESIGN.md section 6).\n\n")
        hdr = ["kernel", "regs", "stack", "smem B", "local", "SASS instr"] + [p[0] for p in PATTERNS]
        f.write("| " + " | ".join(hdr) + " |\n|" + "---|" * len(hdr) + "\n")
        for name, r, st, sh, lo, tot, c in rows:
            f.write("| `" + name + "` | " + " | ".join(str(x) for x in (r, st, sh, lo, tot)) + " | " + " | ".join(str(c.get(p[0], 0)) for p in PATTERNS) + " |\n")
    print(f"{len(rows)} kernels -> profiles/r02_sass_summary.md")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Per-kernel SASS instruction histogram of the shipped libsdrgpu.so (cuobjdump -sass): which kernels use the TMA engine
(UBLKCP = cp.async.bulk, UTMALDG / UTMASTG = tensor-map TMA), packed f32x2 arithmetic (FADD2 / FMUL2 / FFMA2), MUFU.LG2,
and that no tensor-core or cuFFT code is present.  Writes profiles/sass_summary.md.  Runs without a GPU."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sdrainer_b200", "libsdrgpu.so")
COLS = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FADD2", "FMUL2", "FFMA2", "MUFU.LG2", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "DFMA", "HMMA", "UTC"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for c in COLS:
                if op == c or op.startswith(c + ".") or (c in ("UTC", "HMMA", "SYNCS", "BAR") and op.startswith(c)):
                    kernels[cur][c] += 1
    lines = ["# SASS summary of sdrainer_b200/libsdrgpu.so (tools/sass_summary.py; cuobjdump -sass, sm_100a)", "",
             "Counts are static instructions per kernel.  UBLKCP = `cp.async.bulk` (TMA bulk copy), UTMALDG / UTMASTG = tensor-map TMA "
             "load / store, SYNCS = mbarrier operations, FADD2 / FMUL2 / FFMA2 = packed f32x2 arithmetic, HMMA / UTC* = tensor-core "
             "instructions (none: the path is memory-bound by design).", "",
             "| kernel | instr | " + " | ".join(COLS) + " |", "|---|---|" + "---|" * len(COLS)]

    def short(name):
        name = name.replace("(anonymous namespace)::", "").replace("sdr::", "")
        return re.sub(r"\(.*$", "", name)

    for k, c in kernels.items():
        lines.append(f"| `{short(k)}` | {c['_total']} | " + " | ".join(str(c[x]) if c[x] else "" for x in COLS) + " |")
    path = os.path.join(ROOT, "profiles", "sass_summary.md")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print(path, len(kernels), "kernels")


if __name__ == "__main__":
    sys.exit(main())

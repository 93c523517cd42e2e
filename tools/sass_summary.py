#!/usr/bin/env python
"""Per-kernel SASS evidence of the Blackwell-native instructions (B200_PROFILING.md: tcgen05 = UTCHMMA/UTCBAR/LDTM/STTM,
TMA = UTMALDG/UTMASTG/UBLKCP, packed FP32 = FFMA2) in the built library.

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "dpdfnet_b200" / "lib" / "libdpdfnet_b200.so"
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "SYNCS", "FFMA2", "FFMA", "MUFU", "LDGSTS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = kernels.setdefault(re.sub(r"\(.*", "", name).replace("dpdf::", ""), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            cur["_total"] += 1
            if op in MNEMONICS:
                cur[op] += 1
    print(f"# {LIB.name}: SASS instruction counts per kernel (cuobjdump -sass, sm_100a)")
    print(f"{'kernel':44s} {'instrs':>7s} " + " ".join(f"{m:>8s}" for m in MNEMONICS))
    tot = collections.Counter()
    for k, c in kernels.items():
        print(f"{k[:44]:44s} {c['_total']:7d} " + " ".join(f"{c[m]:8d}" for m in MNEMONICS))
        tot.update(c)
    print(f"{'TOTAL':44s} {tot['_total']:7d} " + " ".join(f"{tot[m]:8d}" for m in MNEMONICS))


if __name__ == "__main__":
    sys.exit(main())

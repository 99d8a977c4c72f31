"""Per-kernel counts of the SASS mnemonics that show which hardware path a kernel uses, from `cuobjdump -sass` of the built
library -> profiles/sass_markers.txt.  UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / st (tensor memory),
UBLKCP = cp.async.bulk (TMA engine, non-tensor), UTMALDG / UTMASTG = TMA tensor copies, UTCBAR = tcgen05.commit,
HMMA = mma.sync, LDGSTS = cp.async, SYNCS = mbarrier, MUFU = special-function unit, FFMA = fp32 FMA pipe.
usage: python tools/sass_markers.py [path/to/lib.so]"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARKERS = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "LDGSTS", "SYNCS", "LDSM", "MUFU", "FFMA",
           "DFMA", "DADD", "DMUL", "ATOMS", "SHFL", "VOTE", "BAR")


def main(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
    kernels, cur = OrderedDict(), None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1).split(".")[0]
            kernels[cur]["_total"] += 1
            if op in MARKERS:
                kernels[cur][op] += 1
    try:
        names = subprocess.run(["c++filt"] + list(kernels), check=True, capture_output=True, text=True).stdout.splitlines()
    except Exception:
        names = list(kernels)
    lines = [f"# cuobjdump -sass {os.path.relpath(lib, ROOT)} — instruction counts per kernel (static SASS, sm_100a)",
             "# regenerate: python tools/sass_markers.py"]
    for (mangled, c), name in zip(kernels.items(), names):
        name = re.sub(r"\(.*\)$", "", name).replace("void ", "")
        marks = "  ".join(f"{k}={c[k]}" for k in MARKERS if c[k])
        lines.append(f"{name:70s} total={c['_total']:6d}  {marks}")
    txt = "\n".join(lines) + "\n"
    with open(os.path.join(ROOT, "profiles", "sass_markers.txt"), "w") as f:
        f.write(txt)
    print(txt)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "com_marl_b200", "lib", "libcommarl_b200.so"))

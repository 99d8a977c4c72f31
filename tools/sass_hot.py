"""Summarise a SASS-level ncu source page (tools/sass_capture.sh): opcode mix, stall reasons, blocks of 100 instructions.
usage: python tools/sass_hot.py gpurun_out/TAG_src.csv [units_per_launch]   (units: warps x tiles or envs, for the per-unit figure)"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    units = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, IndexError):
            return 0.0

    def opc(r):
        p = r[ix["Source"]].split()
        o = p[0] if not p[0].startswith("@") else p[1]
        return o.split(".")[0]

    tot = sum(f(r, "Instructions Executed") for r in data)
    thr = sum(f(r, "Thread Instructions Executed") for r in data)
    ts = sum(f(r, "# Samples") for r in data)
    print(rows[0][1])
    print("warp instructions executed: %.0f%s; threads per instruction %.1f; stall samples %d"
          % (tot, " = %.0f per unit" % (tot / units) if units else "", thr / max(tot, 1), ts))
    op, ops = collections.Counter(), collections.Counter()
    for r in data:
        op[opc(r)] += f(r, "Instructions Executed")
        ops[opc(r)] += f(r, "# Samples")
    print("\nopcode      instructions  samples")
    for o, c in op.most_common(24):
        print("%-10s  %5.1f %%       %5.1f %%" % (o, 100 * c / tot, 100 * ops[o] / ts))
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tots = {h: sum(f(r, h) for r in data) for h in st}
    print("\nstall reasons (samples): " + ", ".join("%s %d" % (k[6:], v) for k, v in sorted(tots.items(), key=lambda x: -x[1]) if v > 0))
    print("\nblocks of 100 SASS instructions: share of executed instructions / of samples / barrier-stall samples / threads per instruction / dominant opcodes")
    for b in range(0, len(data), 100):
        blk = data[b:b + 100]
        c = sum(f(r, "Instructions Executed") for r in blk)
        if c == 0:
            continue
        s_ = sum(f(r, "# Samples") for r in blk)
        sb = sum(f(r, "stall_barrier") for r in blk)
        t_ = sum(f(r, "Thread Instructions Executed") for r in blk)
        o2 = collections.Counter()
        for r in blk:
            o2[opc(r)] += f(r, "Instructions Executed")
        print("%5d  %5.2f %%  %5.2f %%  %4d  %4.1f  %s" % (b, 100 * c / tot, 100 * s_ / ts, sb, t_ / c,
                                                      " ".join("%s:%.0f%%" % (o, 100 * v / c) for o, v in o2.most_common(4))))


if __name__ == "__main__":
    main()

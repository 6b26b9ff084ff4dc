"""Top stalled SASS instructions of one launch in an .ncu-rep (source page)."""
import csv
import io
import subprocess
import sys


def main(path, skip=0, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:layer_",
                          "--launch-skip", str(skip), "--launch-count", "1"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    print(rows[0][:2])
    hdr = rows[1]
    isrc, isamp, iexe = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = []
    for r in rows[2:]:
        if len(r) != len(hdr) or not r[isamp].isdigit():
            if data:
                break          # first section = SASS view
            continue
        data.append(r)
    stallcols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp]) for r in data)
    totexe = sum(int(r[iexe]) for r in data)
    print("total samples", tot, "total warp-instructions", totexe)
    agg = {}
    for r in data:
        for i, h in stallcols:
            agg[h] = agg.get(h, 0) + int(r[i] or 0)
    print("stall totals:", sorted(agg.items(), key=lambda t: -t[1])[:8])
    ops = {}
    for r in data:
        op = r[isrc].strip().split()[0] if r[isrc].strip() else "?"
        if op.startswith("@"):
            op = r[isrc].strip().split()[1]
        op = op.split(".")[0]
        o = ops.setdefault(op, [0, 0])
        o[0] += int(r[iexe])
        o[1] += int(r[isamp])
    print("by opcode (executed, samples):")
    for op, (e, s) in sorted(ops.items(), key=lambda t: -t[1][0])[:18]:
        print("   %-10s %12d %5.1f%%   samples %5.1f%%" % (op, e, 100.0 * e / totexe, 100.0 * s / tot))
    for r in sorted(data, key=lambda r: -int(r[isamp]))[:top]:
        st = sorted([(h[6:], int(r[i])) for i, h in stallcols if r[i] not in ("", "0")], key=lambda t: -t[1])[:3]
        print("%6d %5.1f%% exe=%9s  %-64s %s" % (int(r[isamp]), 100 * int(r[isamp]) / tot, r[iexe], r[isrc].strip()[:64], st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40)

"""Stall-reason samples per SASS instruction class / source line of one launch in an .ncu-rep (source page)."""
import csv
import subprocess
import sys
from collections import defaultdict


def main(path, kernel="regex:layer_|front_"):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", kernel,
                          "--launch-count", "1"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = None
    tot = defaultdict(int)
    by_op = defaultdict(lambda: defaultdict(int))
    exe = defaultdict(int)
    for r in rows:
        if r and r[0] in ("Address", "Line No") and "# Samples" in r:
            hdr = r
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            i_src = hdr.index("Source")
            i_exe = hdr.index("Instructions Executed")
            continue
        if hdr and len(r) == len(hdr):
            op = r[i_src].strip().split()
            if not op:
                continue
            name = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
            name = name.split(".")[0] + ("." + name.split(".")[1] if name.startswith(("MUFU", "SYNCS", "LDS", "STG", "LDG", "BAR")) and "." in name else "")
            exe[name] += int(r[i_exe] or 0)
            for i, h in stall_cols:
                v = int(r[i] or 0)
                tot[h] += v
                by_op[name][h] += v
    all_s = sum(tot.values()) or 1
    print("stall samples by reason:", ", ".join("%s=%.1f%%" % (h[6:], 100.0 * v / all_s) for h, v in sorted(tot.items(), key=lambda t: -t[1])[:10]))
    tot_exe = sum(exe.values()) or 1
    print("%-22s %7s %7s  top stall reasons at this instruction" % ("instruction", "exe%", "smp%"))
    for name, d in sorted(by_op.items(), key=lambda t: -sum(t[1].values()))[:28]:
        s = sum(d.values())
        print("%-22s %6.1f%% %6.1f%%  %s" % (name, 100.0 * exe[name] / tot_exe, 100.0 * s / all_s,
                                           ", ".join("%s=%.0f%%" % (h[6:], 100.0 * v / max(s, 1)) for h, v in sorted(d.items(), key=lambda t: -t[1])[:3])))


if __name__ == "__main__":
    main(*sys.argv[1:])

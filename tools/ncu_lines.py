"""Executed instructions and stall samples per CUDA source line of one launch in an .ncu-rep."""
import csv
import subprocess
import sys


def main(path, top=40, kernel="regex:layer_|front_"):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel,
                          "--launch-count", "1"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    fname, hdr, agg = "", None, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 8 and r[0] == "Line No":
            hdr = r
            i_e, i_n = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr and len(r) == len(hdr) and r[0].isdigit() and r[2] == "-":      # a source line row (aggregate, no SASS address)
            key = (fname, int(r[0]))
            e = agg.setdefault(key, [0, 0, r[1].strip()])
            e[0] += int(r[i_e] or 0)
            e[1] += int(r[i_n] or 0)
    tot_e = sum(v[0] for v in agg.values()) or 1
    tot_n = sum(v[1] for v in agg.values()) or 1
    print("total warp-instructions %d, samples %d" % (tot_e, tot_n))
    for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-13s %4d  exe %5.1f%%  samples %5.1f%%  %s" % (f, ln, 100.0 * v[0] / tot_e, 100.0 * v[1] / tot_n, v[2][:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)

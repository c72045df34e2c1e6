"""Developer tool: aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line.

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv;  python tools/analysis/ncu_by_line.py x.csv [rows]

Prints, per source line, the share of warp instructions executed (and instructions per row per warp when `rows` is
given) next to the stall samples - what the fit / QSS kernel notes under profiles/ are read from."""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    hdr = rows[hi]
    iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    cur, curfile = None, None
    instr, samp, txt = collections.Counter(), collections.Counter(), {}
    for r in rows:
        if r and r[0] == "File Path":
            curfile = r[1].split("/")[-1]
            continue
        if len(r) < 10 or r[0] == "Line No":
            continue
        if r[0] != "":
            cur = (curfile, int(r[0]))
            txt[cur] = r[1].strip()[:90]
            continue
        try:
            instr[cur] += float(r[iE])
            samp[cur] += float(r[iS])
        except ValueError:
            pass
    return instr, samp, txt


if __name__ == "__main__":
    instr, samp, txt = load(sys.argv[1])
    nrow = float(sys.argv[2]) if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    tot, tots = sum(instr.values()), sum(samp.values())
    print("warp instructions %.4g, stall samples %d" % (tot, tots))
    for k, v in sorted(samp.items(), key=lambda kv: -kv[1])[:top]:
        extra = "  %6.1f instr/row" % (instr[k] / nrow) if nrow else ""
        print("%-22s samples %5.1f %%  instr %5.1f %%%s  %s" % ("%s:%d" % k, 100 * v / tots, 100 * instr[k] / tot, extra, txt[k]))

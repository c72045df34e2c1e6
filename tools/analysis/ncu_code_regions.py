"""Developer tool: code footprint and time of the QSS kernel by walker, from an ncu source-page export.

    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > x.csv;  python tools/analysis/ncu_code_regions.py x.csv

SASS instructions are walked in address order and attributed to the last walker-level function of sto_qss_memo.cuh whose
lines they carry (inlined helpers inherit the region they were inlined into).  Prints static size, executed size (the part
that competes for the SM's ~32 KB instruction cache), share of executed warp instructions and of stall samples - the
numbers behind the instruction-fetch section of profiles/r01_qss_memo_s4_ncu_summary.txt."""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, "spline_trajectory_optimization_b200", "csrc", "sto_qss_memo.cuh")
WALKERS = {"memo_original_rows", "memo_bwd_rows_group", "memo_spawned_rows", "memo_spawned_rows_group",
           "memo_spawned_rows_vec", "memo_forward_sweep0", "qss_finish_group", "qss_memo_candidate"}


def functions():
    out = []
    for i, line in enumerate(open(SRC), 1):
        m = re.match(r"^(?:STO_HD|STO_D)\s+[\w:<>\*& ]+?\s+(\w+)\s*\(", line)
        if m:
            out.append((i, m.group(1)))
    return out


def main(path):
    funcs = functions()

    def func_of(line):
        name = None
        for n, f in funcs:
            if n <= line:
                name = f
            else:
                break
        return name

    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    hdr = rows[hi]
    iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
    cur, curfile, sass = None, None, []
    for r in rows:
        if r and r[0] == "File Path":
            curfile = r[1].split("/")[-1]
            continue
        if len(r) < 10 or r[0] == "Line No":
            continue
        if r[0] != "":
            cur = (curfile, int(r[0]))
            continue
        try:
            sass.append((int(r[2], 16), cur, float(r[iS]), float(r[iE])))
        except ValueError:
            pass
    sass.sort()
    region = "set-up / slow paths"
    st, hot, ex, sm = (collections.Counter() for _ in range(4))
    for _, c, s, e in sass:
        if c and c[0] == "sto_qss_memo.cuh":
            f = func_of(c[1])
            if f in WALKERS:
                region = f
        st[region] += 1
        ex[region] += e
        sm[region] += s
        hot[region] += 1 if e > 0 else 0
    tot, tots = sum(ex.values()), sum(sm.values())
    print("kernel: %d SASS instructions = %.1f KB, %.1f KB executed at least once" %
          (len(sass), len(sass) * 16 / 1024, sum(hot.values()) * 16 / 1024))
    for k, v in st.most_common():
        print("%-26s static %6.1f KB  executed %6.1f KB  instructions %5.1f %%  stall samples %5.1f %%" %
              (k, v * 16 / 1024, hot[k] * 16 / 1024, 100 * ex[k] / max(tot, 1), 100 * sm[k] / max(tots, 1)))


if __name__ == "__main__":
    main(sys.argv[1])

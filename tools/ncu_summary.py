"""Developer tool: text summary of `ncu --set full` captures (+ the traffic record bench.py keys by kernel-source hash).

    python tools/ncu_summary.py profiles/rNN_name gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...]

Writes profiles/rNN_name_ncu_summary.txt (one block per captured kernel: duration, DRAM bytes, cache hit rates, issue
utilisation, stall mix, instruction-cache hit rate, registers / shared memory) and profiles/rNN_name_traffic.json
({"csrc_sha16": sha of csrc/ in this tree, "kernels": {name: {"dram_bytes_per_launch": ...}}})."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "IPC per SM"), ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "FP64 pipe active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per instruction"),
        ("sm__icc_request_hit_rate.pct", "SM instruction-cache hit %"), ("sm__cycles_elapsed.max", "SM cycles")]


def rows_of(rep):
    """rep: a .ncu-rep, or the `ncu -i x.ncu-rep --page raw --csv` export of one (taken on the GPU box when the report
    itself is too large to bring back)."""
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(out)))
    hdr, units = r[0], r[1]
    return hdr, units, r[2:]


def main():
    prefix, reps = sys.argv[1], sys.argv[2:]
    text, traffic = [], {}
    for rep in reps:
        hdr, units, rows = rows_of(rep)
        for row in rows:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))
            name = re.sub(r"\(.*", "", d.get("Kernel Name", "?"))
            name = re.sub(r"^.*::", "", name)
            short = re.sub(r"<.*", "", name)
            text.append(f"== {d.get('Kernel Name', '?')[:110]}   [{os.path.basename(rep)}]")
            for k, label in KEYS:
                if k in d:
                    text.append(f"   {label:34s} {d[k]} {u.get(k, '')}")
            st = {k: float(d[k].replace(",", "")) for k in hdr
                  if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and d[k]}
            tot = sum(st.values()) or 1.0
            mix = sorted(st.items(), key=lambda x: -x[1])[:8]
            text.append("   stall mix (warp states per issue):  " + ", ".join(
                f"{k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {100 * v / tot:.1f} %"
                for k, v in mix))
            try:
                def to_bytes(key):
                    v, un = float(d[key].replace(",", "")), u.get(key, "").lower()
                    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(un, 1)
                traffic[short] = {"dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
                                  "duration_under_ncu_ms": d.get("gpu__time_duration.sum"), "capture": os.path.basename(rep)}
            except (KeyError, ValueError):
                pass
            text.append("")
    with open(prefix + "_ncu_summary.txt", "w") as f:
        f.write("\n".join(text) + "\n")
    with open(prefix + "_traffic.json", "w") as f:
        json.dump({"csrc_sha16": bench.kernel_source_sha16(), "kernels": traffic}, f, indent=1)
    print("\n".join(text))


if __name__ == "__main__":
    main()

"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on):
   python scripts/ncu_lines.py report.ncu-rep [kernel-regex] -> top lines by instructions and by stall samples."""
import csv
import subprocess
import sys


def main(rep, top=32, extra=()):
    """extra: ncu filter options for reports that hold several launches, e.g. --launch-skip 1 --launch-count 1"""
    raw = subprocess.run(["ncu", "-i", rep, *extra, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    want = ["gpu__time_duration.sum", "launch__grid_size", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.per_cycle_active",
            "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
    for r in rows[2:3]:
        for w in want:
            if w in hdr:
                print(f"{w:70s} {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and "per_issue_active" in h and float(r[i] or 0) > 0.3:
                print("  stall", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), r[i])
    src = subprocess.run(["ncu", "-i", rep, *extra, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                         text=True).stdout.splitlines()
    # one section per source file of the kernel (the kernels live in several included .cuh files)
    idx = [i for i, l in enumerate(src) if l.startswith('"Line No","Source","Address"')]
    names = [l for l in src if l.startswith('"File Name"')]
    agg = {}
    for n, start in enumerate(idx):
        end = idx[n + 1] - 2 if n + 1 < len(idx) else len(src)
        fname = ""
        for back in range(start, max(start - 4, -1), -1):
            if src[back].startswith('"File Name"') or src[back].startswith('"File Path"'):
                fname = src[back].split(",", 1)[1].strip('"').split("/")[-1]
                break
        rows = list(csv.reader(src[start:end]))
        h = rows[0]
        iS, iI, iT = h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
        st = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
        for r in rows[1:]:
            try:
                ln = int(r[0])
            except (ValueError, IndexError):
                continue
            def num(x):
                try:
                    return int(x)
                except ValueError:
                    return 0
            a = agg.setdefault((fname, ln), [r[1], 0, 0, 0, {}])
            a[1] += num(r[iS]); a[2] += num(r[iI]); a[3] += num(r[iT])
            for i in st:
                v = num(r[i])
                if v:
                    a[4][h[i]] = a[4].get(h[i], 0) + v
    ts, ti = sum(a[1] for a in agg.values()), sum(a[2] for a in agg.values())
    print(f"total warp instructions {ti}, samples {ts}")
    per_file = {}
    for (fn, ln), a in agg.items():
        pf = per_file.setdefault(fn, [0, 0, 0])
        pf[0] += a[2]; pf[1] += a[1]; pf[2] += a[3]
    for fn, pf in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        print(f"  file {fn:28s} inst={100 * pf[0] / max(ti, 1):5.1f}% samp={100 * pf[1] / max(ts, 1):5.1f}% thread-inst={pf[2]}")
    keys = set(k for k, _ in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]) | set(k for k, _ in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top])
    for ln in sorted(keys):
        a = agg[ln]
        ss = ",".join(f"{k.replace('stall_', '')}:{v}" for k, v in sorted(a[4].items(), key=lambda kv: -kv[1])[:2])
        print(f"{ln[0][:20]:20s}{ln[1]:5d} inst={100 * a[2] / ti:5.1f}% samp={100 * a[1] / max(ts, 1):5.1f}% lanes={a[3] / max(a[2], 1):4.1f} [{ss}] | {a[0].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], extra=tuple(sys.argv[2:]))

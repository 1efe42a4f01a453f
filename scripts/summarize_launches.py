"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: mean time and share per kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0]
        name = name[5:] if name.startswith("void ") else name   # template instantiations are listed as "void k_x<..>"
        d.setdefault(name, []).append(float(r[vi].replace(",", "")) / 1000.0)
    ours = {k: v for k, v in d.items() if k.startswith("k_")}
    most = max(len(v) for v in ours.values())
    ours = {k: v for k, v in ours.items() if 4 * len(v) >= most}   # per-step kernels only (set-up kernels run once)
    tot = sum(sum(v) / len(v) for v in ours.values())
    for k, v in d.items():
        m = sum(v) / len(v)
        share = f"share={100 * m / tot:5.1f}%" if k in ours else "(set-up / not ours: outside the step)"
        print(f"{k[:44]:44s} n={len(v):4d} mean={m:8.2f} us min={min(v):7.2f} max={max(v):7.2f} {share}")
    print(f"sum of our kernels per step = {tot:.1f} us over {len(ours)} launches")


if __name__ == "__main__":
    main(sys.argv[1])

#!/usr/bin/env python3
"""dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel captured in an `ncu --set full` report, written to
profiles/r2_traffic.json under the key bench.py looks up ("<workload>/<mode>/<state>/<kernel>"):
   python scripts/ncu_traffic.py [--sum [--div N]] G/exact/assembled/scoring gpurun_out/r2_score_G_true.ncu-rep [key report ...]"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def traffic(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    tot, launches = 0.0, 0
    for r in rows[2:]:
        launches += 1
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(r[i]) * UNIT[units[i]]
    name = rows[2][hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    return (tot if SUM else tot / max(launches, 1)), name, launches


SUM = False


def main(argv):
    global SUM
    div = 1.0
    if argv and argv[0] == "--sum":   # the report holds the kernels of ONE evaluation (e.g. k_lnz_refresh + k_lnz_stream): add them up
        SUM, argv = True, argv[1:]
    if argv and argv[0] == "--div":   # ... or of N steps / evaluations: --sum --div N = bytes per step
        div, argv = float(argv[1]), argv[2:]
    path = os.path.join(ROOT, "profiles", "r2_traffic.json")
    d = json.load(open(path)) if os.path.exists(path) else {}
    for key, rep in zip(argv[0::2], argv[1::2]):
        t, name, n = traffic(rep)
        t /= div
        d[key] = t
        d.setdefault("_source", {})[key] = "%s: %s, %d launch(es)" % (os.path.basename(rep), name.split("(")[0], n)
        print(key, t, name.split("(")[0])
    json.dump(d, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])

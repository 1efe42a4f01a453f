#!/usr/bin/env python3
"""N2 measurement: one x3 sub-sampling step of a pyramid level (subsample_data_set) -- the reference's own function (verbatim
copy under baseline/_ref, pure Python dictionaries) against instagraal_b200.pyramid_build (GPU binning) on the same synthetic
level; the outputs must be byte-identical.   python scripts/gpu_r2_pyramid.py --frags 60000 --contacts 3000000"""
import argparse, filecmp, json, os, sys, tempfile, time, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ap = argparse.ArgumentParser()
ap.add_argument("--frags", type=int, default=60000); ap.add_argument("--contacts", type=int, default=3000000)
ap.add_argument("--skip-reference", action="store_true")
a = ap.parse_args()
from instagraal_b200 import pyramid_build as pb
rng = np.random.RandomState(1)
tmp = tempfile.mkdtemp()
n_ctg = 40
per = np.bincount(rng.randint(0, n_ctg, a.frags), minlength=n_ctg) + 1
with open(os.path.join(tmp, "c.txt"), "w") as fc, open(os.path.join(tmp, "f.txt"), "w") as ff:
    fc.write("contig\tlength_kb\tn_frags\tcumul_length\n")
    ff.write("id\tchrom\tstart_pos\tend_pos\tsize\tgc_content\taccu_frag\tfrag_start\tfrag_end\n")
    cum = 0
    for c, n in enumerate(per):
        sizes = rng.randint(100, 4000, n); ends = np.cumsum(sizes); starts = ends - sizes
        fc.write("ctg%d\t%d\t%d\t%d\n" % (c, ends[-1], n, cum))
        gcs = rng.randint(0, 1000, n) / 997.0
        ff.write("".join("%d\tctg%d\t%d\t%d\t%d\t%s\t1\t%d\t%d\n" % (i + 1, c, starts[i], ends[i], sizes[i], float(gcs[i]), i + 1, i + 1) for i in range(n)))
        cum += n
total = int(per.sum())
fa = rng.randint(0, total, a.contacts)
fb = np.where(rng.rand(a.contacts) < 0.8, np.clip(fa + rng.randint(0, 30, a.contacts), 0, total - 1), rng.randint(0, total, a.contacts))
nc = rng.randint(1, 20, a.contacts)
import pandas as pd
with open(os.path.join(tmp, "a.txt"), "w") as fh:
    fh.write("id_frag_a\tid_frag_b\tn_contact\n")
    pd.DataFrame({"a": fa, "b": fb, "n": nc}).to_csv(fh, sep="\t", header=False, index=False)
out = {}
def run(fn, tag, **kw):
    t0 = time.perf_counter()
    nf = fn(os.path.join(tmp, "c.txt"), os.path.join(tmp, "f.txt"), 3, os.path.join(tmp, "a.txt"), os.path.join(tmp, tag + "_a.txt"), 1,
            os.path.join(tmp, tag + "_c.txt"), os.path.join(tmp, tag + "_f.txt"), os.path.join(tmp, tag + "_s.txt"), **kw)
    return nf, time.perf_counter() - t0
pb.bin_contacts([0], [0], [1])   # context creation outside the timed region
nf, t_ours = run(pb.subsample_data_set, "ours")
res = {"n_frags": total, "n_contacts": a.contacts, "n_frags_next_level": int(nf), "ours_s": t_ours}
if not a.skip_reference:
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_harness")); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
    import instagraal.pyramid_sparse as PS
    nf2, t_ref = run(PS.subsample_data_set, "ref")
    same = all(filecmp.cmp(os.path.join(tmp, "ours_" + k), os.path.join(tmp, "ref_" + k), shallow=False) for k in ("a.txt", "c.txt", "f.txt", "s.txt"))
    res.update({"reference_python_s": t_ref, "speedup": t_ref / t_ours, "byte_identical": bool(same and nf == nf2)})
print(json.dumps(res))

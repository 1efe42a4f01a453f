"""TEST INFRASTRUCTURE ONLY: golden files for the pyramid build, made by running the UNMODIFIED reference
functions of /root/reference/src/instagraal/pyramid_sparse.py (init_frag_list, subsample_data_set,
fill_sparse_pyramid_level) on a small synthetic `instagraal-pre` output folder.  h5py is not installed here: the module-level
import is satisfied by an empty stand-in and fill_sparse_pyramid_level gets a recording stand-in for the file handle.

   python -m oracle.make_pyramid_golden       (here, where /root/reference exists) -> tests/golden/pyramid/
"""
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden", "pyramid")


def write_input(folder, seed=7, n_frags=(1, 2, 3, 4, 7, 9, 10, 30, 61, 100, 5, 2, 140)):
    """what instagraal-pre writes (pre.py:244-292): fragments_list.txt, info_contigs.txt, abs_fragments_contacts_weighted.txt"""
    rng = np.random.RandomState(seed)
    n_frags = list(n_frags)
    os.makedirs(folder, exist_ok=True)
    with open(os.path.join(folder, "info_contigs.txt"), "w") as fc, open(os.path.join(folder, "fragments_list.txt"), "w") as ff:
        fc.write("contig\tlength\tn_frags\tcumul_length\n")
        ff.write("id\tchrom\tstart_pos\tend_pos\tsize\tgc_content\n")
        cumul = 0
        for c, n in enumerate(n_frags):
            name = "ctg%02d" % c
            sizes = rng.randint(40, 3000, n)
            ends = np.cumsum(sizes)
            starts = ends - sizes
            fc.write("%s\t%d\t%d\t%d\n" % (name, ends[-1], n, cumul))
            for i in range(n):
                gc = [0.5, 0.0, 1.0 / 3.0, rng.randint(0, 1000) / 997.0, round(rng.rand(), 4)][rng.randint(5)]
                ff.write("%d\t%s\t%d\t%d\t%d\t%s\n" % (i + 1, name, starts[i], ends[i], sizes[i], gc))
            cumul += n
    total = cumul
    # contacts: mostly near the diagonal, some far, duplicates, both orders of the pair, unsorted
    m = 6000
    a = rng.randint(0, total, m)
    b = np.where(rng.rand(m) < 0.8, np.clip(a + rng.randint(-6, 7, m), 0, total - 1), rng.randint(0, total, m))
    nc = rng.randint(1, 40, m)
    with open(os.path.join(folder, "abs_fragments_contacts_weighted.txt"), "w") as fh:
        fh.write("%d\t%d\t%d\n" % (total, total, m))
        for i in range(m):
            fh.write("%d\t%d\t%d\n" % (a[i], b[i], nc[i]))
    return total


class _Dataset:
    def __init__(self, shape, dtype):
        self.a = np.zeros(shape, dtype=np.int32)

    def __setitem__(self, k, v):
        self.a[k] = v


class _Group:
    def __init__(self):
        self.d = {}

    def create_dataset(self, name, shape, dtype):
        self.d[name] = _Dataset(shape, dtype)
        return self.d[name]


class FakeH5:
    def __init__(self):
        self.g, self.attrs = {}, {}

    def create_group(self, name):
        self.g[name] = _Group()
        return self.g[name]


def reference_module():
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    sys.path.insert(0, os.path.join(HERE, "ref_harness"))   # empty matplotlib stand-in
    sys.path.insert(0, "/root/reference/src")
    import instagraal.pyramid_sparse as PS
    return PS


def run(PS, base, out, n_levels=4, factor=3, min_bin=1):
    """the level loop of build() (PS:178-277) with the reference's own functions; returns {level: (3, nnz) array, nfrags}"""
    h5 = FakeH5()
    res = {}
    cur_contigs = cur_frags = cur_contacts = s2s = None
    for level in range(n_levels):
        d = os.path.join(out, "level_%d" % level)
        os.makedirs(d, exist_ok=True)
        pre = "%d_" % level
        if level == 0:
            cur_contigs, cur_frags, cur_contacts = (os.path.join(d, pre + x) for x in ("contig_info.txt", "fragments_list.txt", "abs_frag_contacts.txt"))
            shutil.copyfile(os.path.join(base, "info_contigs.txt"), cur_contigs)
            shutil.copyfile(os.path.join(base, "abs_fragments_contacts_weighted.txt"), cur_contacts)
            nfrags = PS.init_frag_list(os.path.join(base, "fragments_list.txt"), cur_frags)
        else:
            nc_, nf_, na_ = (os.path.join(d, pre + x) for x in ("contig_info.txt", "fragments_list.txt", "abs_frag_contacts.txt"))
            nfrags = PS.subsample_data_set(cur_contigs, cur_frags, factor, cur_contacts, na_, min_bin, nc_, nf_, s2s)
            cur_contigs, cur_frags, cur_contacts = nc_, nf_, na_
        PS.fill_sparse_pyramid_level(h5, level, cur_contacts, nfrags)
        res["data_%d" % level] = h5.g[str(level)].d["data"].a.copy()
        res["nfrags_%d" % level] = np.int64(nfrags)
        s2s = os.path.join(d, pre + "sub_2_super_index_frag.txt")
    return res


class _NoPlot:
    """remove_problematic_fragments draws two diagnostic plots into the working directory (PS:764-770); not part of any format"""

    def __getattr__(self, name):
        return lambda *a, **k: None


def run_filter(PS, level0_dir, data0, nfrags0, out_dir, thresh_factor=1):
    """the UNMODIFIED reference remove_problematic_fragments (PS:731-1030) on a level-0 folder; returns the threshold"""
    PS.plt = _NoPlot()
    os.makedirs(out_dir, exist_ok=True)
    src = {k: os.path.join(level0_dir, "0_" + k) for k in ("contig_info.txt", "fragments_list.txt", "abs_frag_contacts.txt")}
    dst = {k: os.path.join(out_dir, "0_" + k) for k in src}
    pyr0 = {"0": {"data": data0, "nfrags": np.array([[int(nfrags0)]], dtype=np.int32)}}
    return PS.remove_problematic_fragments(src["contig_info.txt"], src["fragments_list.txt"], src["abs_frag_contacts.txt"],
                                           dst["contig_info.txt"], dst["fragments_list.txt"], dst["abs_frag_contacts.txt"], pyr0,
                                           thresh_factor=thresh_factor)


if __name__ == "__main__":
    PS = reference_module()
    for sub in (os.listdir(OUT) if os.path.isdir(OUT) else []):   # (load_golden.npz is written by make_pyramid_load_golden)
        if sub in ("input", "expected") or sub.startswith("filtered_"):
            shutil.rmtree(os.path.join(OUT, sub))
    base = os.path.join(OUT, "input")
    write_input(base)
    res = run(PS, base, os.path.join(OUT, "expected"))
    np.savez_compressed(os.path.join(OUT, "expected", "hdf5_arrays.npz"), **res)
    for tf in (1, 0.25):   # the filtering step of build_and_filter on the level-0 files
        fdir = os.path.join(OUT, "filtered_%s" % str(tf).replace(".", "p"))
        th = run_filter(PS, os.path.join(OUT, "expected", "level_0"), res["data_0"], res["nfrags_0"], fdir, thresh_factor=tf)
        with open(os.path.join(fdir, "thresh.txt"), "w") as fh:
            fh.write(repr(float(th)) + "\n")
    for f in os.listdir(ROOT):   # the reference's logger drops a file into the working directory
        if f.startswith("instagraal-") and f.endswith(".log"):
            os.remove(os.path.join(ROOT, f))
    print("written", OUT, {k: (v.shape if hasattr(v, "shape") and v.shape else int(v)) for k, v in res.items()})

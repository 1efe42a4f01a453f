"""Pyramid build: instagraal_b200.pyramid_build against golden files written by the UNMODIFIED reference
functions (oracle/make_pyramid_golden.py ran pyramid_sparse.init_frag_list / subsample_data_set / fill_sparse_pyramid_level on
tests/golden/pyramid/input): every text file of every level byte for byte, the (3, nnz) HDF5 arrays element for element --
including the reference's quirk Q13 (the first data line of a contact file is dropped by subsample_data_set)."""
import filecmp
import os

import numpy as np
import pytest

from instagraal_b200 import pyramid_build as pb

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pyramid")
N_LEVELS = 4


def _paths(root, level):
    d = os.path.join(root, "level_%d" % level)
    pre = "%d_" % level
    return {k: os.path.join(d, pre + k) for k in ("contig_info.txt", "fragments_list.txt", "abs_frag_contacts.txt", "sub_2_super_index_frag.txt")}


def _same(a, b):
    assert filecmp.cmp(a, b, shallow=False), (a, b)


def test_fragment_and_contig_bookkeeping_matches_reference(built, tmp_path):
    """host part (no GPU): level-0 fragment list, then contig list / fragment list / old->new index of levels 1..3"""
    exp0 = _paths(os.path.join(G, "expected"), 0)
    os.makedirs(tmp_path / "level_0")
    got0 = _paths(str(tmp_path), 0)
    n = pb.init_frag_list(os.path.join(G, "input", "fragments_list.txt"), got0["fragments_list.txt"])
    assert n == 374
    _same(got0["fragments_list.txt"], exp0["fragments_list.txt"])
    want_nfrags = np.load(os.path.join(G, "expected", "hdf5_arrays.npz"))
    for level in range(1, N_LEVELS):
        prev, exp = _paths(os.path.join(G, "expected"), level - 1), _paths(os.path.join(G, "expected"), level)
        os.makedirs(tmp_path / ("level_%d" % level))
        got = _paths(str(tmp_path), level)
        s2s = os.path.join(str(tmp_path), "s2s_%d.txt" % level)
        nf = pb.subsample_data_set(prev["contig_info.txt"], prev["fragments_list.txt"], 3, "SIMU", got["abs_frag_contacts.txt"], 1,
                                   got["contig_info.txt"], got["fragments_list.txt"], s2s)
        assert nf == int(want_nfrags["nfrags_%d" % level])
        _same(got["contig_info.txt"], exp["contig_info.txt"])
        _same(got["fragments_list.txt"], exp["fragments_list.txt"])
        _same(s2s, prev["sub_2_super_index_frag.txt"])
    # factor 1: plain copies + identity index (PS:482-495)
    os.makedirs(tmp_path / "f1")
    f1 = {k: str(tmp_path / "f1" / k) for k in ("c", "f", "a", "s")}
    p1 = _paths(os.path.join(G, "expected"), 1)
    assert pb.subsample_data_set(p1["contig_info.txt"], p1["fragments_list.txt"], 1, p1["abs_frag_contacts.txt"], f1["a"], 1, f1["c"], f1["f"], f1["s"]) == 132
    _same(f1["a"], p1["abs_frag_contacts.txt"])
    assert open(f1["s"]).read().splitlines()[:3] == ["current_id\tsuper_id", "1\t1", "2\t2"]


class _Dataset:
    def __init__(self, shape):
        self.a = np.zeros(shape, dtype=np.int32)

    def __setitem__(self, k, v):
        self.a[k] = v


class _Group:
    def __init__(self):
        self.d = {}

    def create_dataset(self, name, shape, dtype):
        self.d[name] = _Dataset(shape)
        return self.d[name]


class FakeH5:
    def __init__(self):
        self.g, self.attrs = {}, {}

    def create_group(self, name):
        self.g[name] = _Group()
        return self.g[name]


@pytest.mark.gpu
def test_build_matches_reference_levels_and_hdf5_arrays(built, tmp_path):
    """the whole level loop through the C ABI (ig_bin_contacts): contact files of every level and the HDF5 group contents"""
    base = os.path.join(G, "input")
    nfr = pb.build(base, N_LEVELS, 3, 1, output_folder=str(tmp_path))
    root = os.path.join(str(tmp_path), "pyramids", "pyramid_%d_no_thresh" % N_LEVELS)
    want = np.load(os.path.join(G, "expected", "hdf5_arrays.npz"))
    assert nfr == [int(want["nfrags_%d" % lv]) for lv in range(N_LEVELS)]
    for level in range(N_LEVELS):
        got, exp = _paths(root, level), _paths(os.path.join(G, "expected"), level)
        for k in ("contig_info.txt", "fragments_list.txt", "abs_frag_contacts.txt"):
            _same(got[k], exp[k])
        if level < N_LEVELS - 1:
            _same(got["sub_2_super_index_frag.txt"], exp["sub_2_super_index_frag.txt"])
        h5 = FakeH5()
        arr = pb.fill_sparse_pyramid_level(h5, level, got["abs_frag_contacts.txt"], nfr[level])
        assert np.array_equal(arr, want["data_%d" % level]), level
        assert np.array_equal(h5.g[str(level)].d["data"].a, want["data_%d" % level])
        assert int(h5.g[str(level)].d["nfrags"].a[0, 0]) == nfr[level]


@pytest.mark.gpu
def test_bin_contacts_edge_cases(built):
    a, b, n = pb.bin_contacts(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32))
    assert len(a) == 0
    # one pair given in both orders and twice: summed once, ordered
    a, b, n = pb.bin_contacts([5, 2, 5, 0], [2, 5, 5, 0], [3, 4, 1, 9])
    assert a.tolist() == [0, 2, 5] and b.tolist() == [0, 5, 5] and n.tolist() == [9, 7, 1]
    # first-appearance order inside a row; sums above int32
    a, b, n = pb.bin_contacts([1, 1, 1, 0, 1], [9, 3, 9, 4, 3], [2**31 - 1, 1, 2**31 - 1, 5, 1], first_appearance_order=True)
    assert a.tolist() == [0, 1, 1] and b.tolist() == [4, 9, 3] and n.tolist() == [5, 2 * (2**31 - 1), 2]
    with pytest.raises(RuntimeError, match="out of range"):
        pb.bin_contacts([0, 7], [1, 1], [1, 1], old2new=[0, 0, 1])
    # a larger random case against NumPy
    rng = np.random.RandomState(1)
    fa, fb, nc = rng.randint(0, 5000, 400000), rng.randint(0, 5000, 400000), rng.randint(1, 9, 400000)
    m = rng.randint(0, 1700, 5000)
    a, b, n = pb.bin_contacts(fa, fb, nc, old2new=m)
    lo, hi = np.minimum(m[fa], m[fb]), np.maximum(m[fa], m[fb])
    key = lo.astype(np.int64) * 1700 + hi
    uk, inv = np.unique(key, return_inverse=True)
    assert np.array_equal(a.astype(np.int64) * 1700 + b, uk)
    assert np.array_equal(n, np.bincount(inv, weights=nc).astype(np.int64))


# ---- the filtering step of build_and_filter (PS:731-1030) ------------------------------------------------------------------
def _numpy_bin_contacts(fa, fb, nc, old2new=None, first_appearance_order=False, device=0):
    """test-local NumPy statement of what ig_bin_contacts returns, so that the HOST logic around it is checked without a GPU"""
    fa, fb, nc = np.asarray(fa, np.int64), np.asarray(fb, np.int64), np.asarray(nc, np.int64)
    if old2new is not None:
        m = np.asarray(old2new, np.int64)
        fa, fb = m[fa], m[fb]
    lo, hi = np.minimum(fa, fb), np.maximum(fa, fb)
    key = lo * (1 << 32) + hi
    uk, first, inv = np.unique(key, return_index=True, return_inverse=True)
    sums = np.bincount(inv, weights=nc, minlength=len(uk)).astype(np.int64)
    order = np.lexsort((first, uk >> 32)) if first_appearance_order else np.arange(len(uk))
    return (uk[order] >> 32).astype(np.int32), (uk[order] & 0xffffffff).astype(np.int32), sums[order]


def _filter_paths(d):
    return {k: os.path.join(d, "0_" + k) for k in ("contig_info.txt", "fragments_list.txt", "abs_frag_contacts.txt")}


def _run_filter(tmp, thresh_factor):
    z = np.load(os.path.join(G, "expected", "hdf5_arrays.npz"))
    src, dst = _filter_paths(os.path.join(G, "expected", "level_0")), _filter_paths(str(tmp))
    pyr0 = {"0": {"data": z["data_0"], "nfrags": np.array([[int(z["nfrags_0"])]], dtype=np.int32)}}
    th = pb.remove_problematic_fragments(src["contig_info.txt"], src["fragments_list.txt"], src["abs_frag_contacts.txt"], dst["contig_info.txt"],
                                         dst["fragments_list.txt"], dst["abs_frag_contacts.txt"], pyr0, thresh_factor=thresh_factor)
    return th, dst


@pytest.mark.parametrize("tf,folder", [(1, "filtered_1"), (0.25, "filtered_0p25")])
def test_filtering_bookkeeping_matches_reference(tmp_path, monkeypatch, tf, folder):
    """host part: locked fragments merged forward, trailing runs and whole contigs destroyed, the accu_frag leak across a contig
    start, GC means, contig list, threshold -- against files written by the reference's own remove_problematic_fragments"""
    monkeypatch.setattr(pb, "bin_contacts", _numpy_bin_contacts)
    th, dst = _run_filter(tmp_path, tf)
    exp = _filter_paths(os.path.join(G, folder))
    assert repr(float(th)) == open(os.path.join(G, folder, "thresh.txt")).read().strip()
    for k in dst:
        _same(dst[k], exp[k])


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/instagraal"), reason="differential run against the live reference (build container only)")
@pytest.mark.parametrize("seed,n_frags,tf", [(21, (30, 1, 2, 50, 3), 1), (22, (5, 5, 5, 5, 80), 0.0), (23, (120,), 0.5), (24, (2, 2, 2, 2, 2, 2, 40), -0.5)])
def test_filtering_differential_against_the_live_reference(tmp_path, monkeypatch, seed, n_frags, tf):
    import subprocess
    import sys
    root = os.path.dirname(G.rstrip("/").rsplit("/tests/", 1)[0] + "/x")
    code = (
        "import sys, os, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import oracle.make_pyramid_golden as MG\n"
        "PS = MG.reference_module()\n"
        "base, out = %r, %r\n"
        "MG.write_input(base, seed=%d, n_frags=%r)\n"
        "res = MG.run(PS, base, out, n_levels=1)\n"
        "np.savez(os.path.join(out, 'h5.npz'), **res)\n"
        "th = MG.run_filter(PS, os.path.join(out, 'level_0'), res['data_0'], res['nfrags_0'], os.path.join(out, 'filt'), thresh_factor=%r)\n"
        "open(os.path.join(out, 'filt', 'thresh.txt'), 'w').write(repr(float(th)))\n"
    ) % (root, str(tmp_path / "in"), str(tmp_path / "pyr"), seed, tuple(n_frags), tf)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    monkeypatch.setattr(pb, "bin_contacts", _numpy_bin_contacts)
    out = str(tmp_path / "pyr")
    z = np.load(os.path.join(out, "h5.npz"))
    src, dst = _filter_paths(os.path.join(out, "level_0")), _filter_paths(str(tmp_path / "mine"))
    os.makedirs(tmp_path / "mine")
    pyr0 = {"0": {"data": z["data_0"], "nfrags": np.array([[int(z["nfrags_0"])]], dtype=np.int32)}}
    th = pb.remove_problematic_fragments(src["contig_info.txt"], src["fragments_list.txt"], src["abs_frag_contacts.txt"], dst["contig_info.txt"],
                                         dst["fragments_list.txt"], dst["abs_frag_contacts.txt"], pyr0, thresh_factor=tf)
    exp = _filter_paths(os.path.join(out, "filt"))
    assert repr(float(th)) == open(os.path.join(out, "filt", "thresh.txt")).read()
    for k in dst:
        _same(dst[k], exp[k])


@pytest.mark.gpu
def test_build_and_filter_end_to_end(built, tmp_path):
    """pre-processing output -> unfiltered level -> filtered level 0 (== the reference's files) -> level loop -> loaded pyramid"""
    pyr = pb.build_and_filter(os.path.join(G, "input"), 3, 3, thresh_factor=1, output_folder=str(tmp_path))
    root = os.path.join(str(tmp_path), "pyramids", "pyramid_3_thresh_auto")
    got, exp = _filter_paths(os.path.join(root, "level_0")), _filter_paths(os.path.join(G, "filtered_1"))
    for k in got:
        _same(got[k], exp[k])
    assert os.path.isdir(os.path.join(str(tmp_path), "pyramids", "pyramid_1_no_thresh", "level_0"))
    n0 = pb.file_len(got["fragments_list.txt"]) - 1
    lev0, lev1 = pyr.get_level(0), pyr.get_level(1)
    assert lev0.n_frags == n0 == 311 and lev0.n_contigs == 11
    assert lev1.n_frags == pb.file_len(os.path.join(root, "level_1", "1_fragments_list.txt")) - 1
    assert int(lev1.S_o_A_frags["sub_len"].sum()) == n0           # every level-0 fragment sits in exactly one level-1 bin
    assert lev1.sparse_mat_csr.sum() <= lev0.sparse_mat_csr.sum()  # (Q13: each binning step drops the first data line)
    # a second call finds everything built and only loads
    pyr2 = pb.build_and_filter(os.path.join(G, "input"), 3, 3, thresh_factor=1, output_folder=str(tmp_path))
    assert pyr2.get_level(1).n_frags == lev1.n_frags


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/instagraal"), reason="differential run against the live reference (build container only)")
@pytest.mark.parametrize("seed,n_frags,factor,min_bin", [(31, (1, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 27, 28), 3, 1), (32, (100, 3, 50), 2, 1),
                                                           (33, (10, 20, 30, 2), 3, 4), (34, (64, 65, 1), 4, 1), (35, (200,), 9, 1)])
def test_level_loop_differential_against_the_live_reference(tmp_path, monkeypatch, seed, n_frags, factor, min_bin):
    """init_frag_list + subsample_data_set + fill_sparse_pyramid_level for 4 levels on fresh inputs, other binning factors
    (>= 8 exercises NumPy's pairwise mean of the GC contents) and min_bin_per_contig > 1 (contigs left unbinned)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, os, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import oracle.make_pyramid_golden as MG\n"
        "PS = MG.reference_module()\n"
        "MG.write_input(%r, seed=%d, n_frags=%r)\n"
        "res = MG.run(PS, %r, %r, n_levels=4, factor=%d, min_bin=%d)\n"
        "np.savez(os.path.join(%r, 'h5.npz'), **res)\n"
    ) % (root, str(tmp_path / "in"), seed, tuple(n_frags), str(tmp_path / "in"), str(tmp_path / "ref"), factor, min_bin, str(tmp_path / "ref"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    monkeypatch.setattr(pb, "bin_contacts", _numpy_bin_contacts)
    want = np.load(os.path.join(str(tmp_path / "ref"), "h5.npz"))
    base, out = str(tmp_path / "in"), str(tmp_path / "mine")
    cur = None
    for level in range(4):
        os.makedirs(os.path.join(out, "level_%d" % level))
        got, exp = _paths(out, level), _paths(str(tmp_path / "ref"), level)
        if level == 0:
            import shutil
            shutil.copyfile(os.path.join(base, "info_contigs.txt"), got["contig_info.txt"])
            shutil.copyfile(os.path.join(base, "abs_fragments_contacts_weighted.txt"), got["abs_frag_contacts.txt"])
            nfr = pb.init_frag_list(os.path.join(base, "fragments_list.txt"), got["fragments_list.txt"])
        else:
            nfr = pb.subsample_data_set(cur["contig_info.txt"], cur["fragments_list.txt"], factor, cur["abs_frag_contacts.txt"],
                                        got["abs_frag_contacts.txt"], min_bin, got["contig_info.txt"], got["fragments_list.txt"],
                                        cur["sub_2_super_index_frag.txt"])
            _same(cur["sub_2_super_index_frag.txt"], _paths(str(tmp_path / "ref"), level - 1)["sub_2_super_index_frag.txt"])
        assert nfr == int(want["nfrags_%d" % level])
        for k in ("contig_info.txt", "fragments_list.txt", "abs_frag_contacts.txt"):
            _same(got[k], exp[k])
        arr = pb.fill_sparse_pyramid_level(None, level, got["abs_frag_contacts.txt"], nfr)
        assert np.array_equal(arr, want["data_%d" % level]), level
        cur = got

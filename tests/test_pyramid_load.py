"""N2, load side: instagraal_b200.pyramid_load.{pyramid, level} against golden vectors dumped from the UNMODIFIED reference
classes (oracle/make_pyramid_load_golden.py: pyramid_sparse.pyramid / level run on tests/golden/pyramid/expected/).  Host code:
runs without a GPU (the contact arrays are served from the recorded HDF5 arrays, as h5py would)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR

FOLDER = os.path.join(GOLDEN_DIR, "pyramid", "expected")
N_LEVELS = 4
SOA_KEYS = ["pos", "sub_pos", "id_c", "start_bp", "len_bp", "sub_len", "circ", "id", "prev", "next", "l_cont", "sub_l_cont",
            "l_cont_bp", "n_accu"]
FRAG_INFO_KEYS = ["index", "start_pos(bp)", "end_pos(bp)", "size(bp)", "sub_low_index", "sub_high_index", "super_index", "n_accu_frags"]
B_FRAG_INT = ["id_init", "start_pos", "end_pos", "length_kb", "np_id_abs", "curr_id", "contig_id", "init_frag_start", "init_frag_end",
              "sub_frag_start", "sub_frag_end", "super_index", "n_accu_frags"]


class RecordedHdf5:
    def __init__(self):
        z = np.load(os.path.join(FOLDER, "hdf5_arrays.npz"))
        self.g = {}
        for k in z.files:
            kind, lvl = k.rsplit("_", 1)
            self.g.setdefault(lvl, {})[kind] = z[k] if kind == "data" else np.array([[int(z[k])]], dtype=np.int32)

    def __getitem__(self, k):
        return self.g[k]

    def close(self):
        pass


@pytest.fixture(scope="module")
def loaded():
    from instagraal_b200.pyramid_load import pyramid
    return pyramid(FOLDER, N_LEVELS, data=RecordedHdf5()), np.load(os.path.join(GOLDEN_DIR, "pyramid", "load_golden.npz"))


def test_pyramid_dictionaries_match_the_reference(loaded):
    pyr, g = loaded
    assert list(pyr.list_contigs_name) == list(g["list_contigs_name"])
    assert list(pyr.list_contigs_id) == list(g["list_contigs_id"])
    for l in range(N_LEVELS):
        sp = pyr.spec_level[str(l)]
        fd = sp["fragments_dict"]
        ids = sorted(fd.keys())
        assert ids == list(g["L%d_fd_ids" % l])
        for k in FRAG_INFO_KEYS:
            assert [fd[i][k] for i in ids] == list(g["L%d_fd_%s" % (l, k)]), (l, k)
        assert [fd[i]["init_contig"] for i in ids] == list(g["L%d_fd_init_contig" % l])
        assert [fd[i]["tag"] for i in ids] == list(g["L%d_fd_tag" % l])
        cd = sp["contigs_dict"]
        assert [str(k) for k in cd.keys()] == list(g["L%d_cd_keys" % l])
        frs = [f for k in cd.keys() if isinstance(k, int) for f in cd[k]]
        for k in B_FRAG_INT:
            assert [getattr(f, k) for f in frs] == list(g["L%d_bf_%s" % (l, k)]), (l, k)
        assert np.array_equal(np.array([f.gc_content for f in frs]), g["L%d_bf_gc" % l])
        assert [f.init_name for f in frs] == list(g["L%d_bf_init_name" % l])
        assert all(f.orientation == "w" and f.curr_name == "" and f.pos_kb == 0 for f in frs)


@pytest.mark.parametrize("l", range(N_LEVELS))
def test_level_load_data_matches_the_reference(loaded, l):
    pyr, g = loaded
    lev = pyr.get_level(l)
    assert lev.n_frags == int(g["L%d_n_frags" % l]) and lev.n_contigs == int(g["L%d_n_contigs" % l])
    assert type(lev.mean_value_trans).__name__ == str(g["L%d_mean_value_trans_type" % l])
    assert float(lev.mean_value_trans) == float(g["L%d_mean_value_trans" % l])          # bit for bit
    for k in SOA_KEYS:
        assert lev.S_o_A_frags[k].dtype == np.int32
        assert np.array_equal(lev.S_o_A_frags[k], g["L%d_soa_%s" % (l, k)]), k
    assert list(lev.S_o_A_frags.keys()) == SOA_KEYS
    assert np.array_equal(np.array([list(t) for t in lev.vect_frag_np.tolist()], dtype=np.int32).reshape(-1, 11), g["L%d_vect_frag_np" % l])
    assert lev.vect_frag_np.dtype.names == ("pos", "id_c", "start_bp", "len_bp", "circ", "id", "prev", "next", "l_cont", "l_cont_bp", "n_accu")
    assert np.array_equal(lev.distri_frag, g["L%d_distri_frag" % l])
    assert list(lev.frags_init_contigs) == list(g["L%d_frags_init_contigs" % l])
    assert np.array_equal(lev.pos_vect_frags_4_GL, g["L%d_pos_gl" % l])
    assert np.array_equal(lev.col_vect_frags_4_GL, g["L%d_col_gl" % l])
    for nm, m in (("csr", lev.sparse_mat_csr), ("csc", lev.sparse_mat_csc)):
        assert np.array_equal(m.data, g["L%d_%s_data" % (l, nm)])
        assert np.array_equal(m.indices, g["L%d_%s_indices" % (l, nm)])
        assert np.array_equal(m.indptr, g["L%d_%s_indptr" % (l, nm)])
    assert list(lev.dict_contigs.keys()) == list(pyr.list_contigs_id)
    for c in lev.dict_contigs:
        dc = lev.dict_contigs[c]
        assert list(dc["intra_coord"]) == list(g["L%d_c%d_intra_coord" % (l, c)])
        assert np.array_equal(dc["tick_kb"], g["L%d_c%d_tick_kb" % (l, c)]) and dc["tick_kb"].dtype == g["L%d_c%d_tick_kb" % (l, c)].dtype
        assert np.array_equal(dc["end_frags_kb"], g["L%d_c%d_end_frags_kb" % (l, c)])
        assert dc["name"] == str(g["L%d_c%d_name" % (l, c)])
        assert dc["frags"] is pyr.spec_level[str(l)]["contigs_dict"][c]


def test_zoom_between_levels(loaded):
    pyr, g = loaded
    fd2 = pyr.spec_level["2"]["fragments_dict"]
    for frag in (1, 7, len(fd2)):
        subs = pyr.zoom_in_frag((frag, 2))
        assert subs == [(i, 1) for i in range(fd2[frag]["sub_low_index"], fd2[frag]["sub_high_index"] + 1)]
        assert pyr.zoom_out_frag((frag, 2)) == (fd2[frag]["super_index"], 3)
        assert pyr.zoom_in_frag((frag, 0)) == [(frag, 0)]
    px = pyr.zoom_in_pixel([2, 5, 2])
    assert px == [min(fd2[2]["sub_low_index"], fd2[5]["sub_low_index"]), max(fd2[2]["sub_high_index"], fd2[5]["sub_high_index"]), 1]
    assert pyr.zoom_in_area([[2, 5, 2], [3, 9, 2]])[0][2] == 1


def test_sequences_and_export_through_the_level(loaded, tmp_path):
    """build_seq_per_bin / load_reference_sequence (incl. the reference's dropped last line) and generate_new_fasta wired
    through level (PS:1938-2033)"""
    pyr, g = loaded
    lev = pyr.get_level(2)
    rng = np.random.RandomState(3)
    fa = tmp_path / "genome.fa"
    seqs = {}
    with open(fa, "w") as fh:
        for c, nm in enumerate(pyr.list_contigs_name):
            ln = int(max(f.end_pos for f in pyr.spec_level["0"]["contigs_dict"][c + 1]))
            s = "".join(rng.choice(list("ACGTacgtN"), ln))
            seqs[nm] = s
            fh.write(">%s some description\n" % nm)
            for i in range(0, ln, 70):
                fh.write(s[i:i + 70] + "\n")
    lev.build_seq_per_bin(str(fa))
    last = pyr.list_contigs_name[-1]
    for nm in pyr.list_contigs_name[:-1]:
        assert pyr.dict_sequence_contigs[nm] == seqs[nm]
    tail = len(seqs[last]) % 70 or 70
    assert pyr.dict_sequence_contigs[last] == seqs[last][:-tail]          # the reference drops the file's last line (PS:1649)
    frs = [f for k in sorted(pyr.spec_level["2"]["contigs_dict"]) for f in pyr.spec_level["2"]["contigs_dict"][k]]
    assert lev.list_seq == [pyr.dict_sequence_contigs[f.init_contig][f.start_pos:f.end_pos] for f in frs]

    class V:
        pass
    v = V()
    v.id_c, v.pos, v.id_d = lev.S_o_A_frags["id_c"].copy(), lev.S_o_A_frags["pos"].copy(), lev.S_o_A_frags["id"].copy()
    v.ori, v.activ = np.ones(lev.n_frags, dtype=np.int32), np.ones(lev.n_frags, dtype=np.int32)
    lev.generate_new_fasta(v, str(tmp_path / "out.fa"), str(tmp_path / "info.txt"))
    txt = open(tmp_path / "out.fa").read()
    assert txt.count(">3C-assembly-contig_") == lev.n_contigs


@pytest.mark.gpu
def test_contacts_rebuilt_from_the_text_levels_when_h5py_is_missing(built):
    """without h5py (this image) and without data=, the (3, nnz) arrays come from level_k/k_abs_frag_contacts.txt through
    the GPU binning of pyramid_build.fill_sparse_pyramid_level: same arrays as the reference's HDF5 writer recorded"""
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py present: the HDF5 cache is read instead")
    except ImportError:
        pass
    from instagraal_b200.pyramid_load import pyramid
    g = np.load(os.path.join(GOLDEN_DIR, "pyramid", "load_golden.npz"))
    z = np.load(os.path.join(FOLDER, "hdf5_arrays.npz"))
    pyr = pyramid(FOLDER, N_LEVELS)
    for l in range(N_LEVELS):
        lev = pyr.get_level(l)
        assert np.array_equal(lev.np_2_scipy_sparse, z["data_%d" % l])
        assert np.array_equal(lev.sparse_mat_csr.data, g["L%d_csr_data" % l])
        assert np.array_equal(lev.sparse_mat_csr.indices, g["L%d_csr_indices" % l])
        assert float(lev.mean_value_trans) == float(g["L%d_mean_value_trans" % l])
    pyr.close()


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/instagraal"), reason="differential run against the live reference (build container only)")
@pytest.mark.parametrize("seed,n_frags", [(11, (40,)), (12, (3, 1, 1, 25, 2)), (13, (9, 9, 9, 9, 9, 9, 9, 1))])
def test_differential_against_the_live_reference_classes(tmp_path, seed, n_frags):
    """fresh pyramid folders written by the reference's own build functions, loaded by the reference's classes and by ours:
    a single contig (NaN -> min/10 fallback of mean_value_trans), one-fragment contigs, equal-length contigs"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, os, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import oracle.make_pyramid_load_golden as ML, oracle.make_pyramid_golden as MG\n"
        "PS = ML.reference_module()\n"
        "base, out = %r, %r\n"
        "MG.write_input(base, seed=%d, n_frags=%r)\n"
        "res = MG.run(PS, base, out)\n"
        "np.savez(os.path.join(out, 'hdf5_arrays.npz'), **res)\n"
        "ML.FOLDER = out\n"
        "import warnings; warnings.simplefilter('ignore')\n"
        "np.savez(os.path.join(out, 'load.npz'), **ML.dump(PS.pyramid(out, ML.N_LEVELS), {}))\n"
    ) % (root, str(tmp_path / "in"), str(tmp_path / "pyr"), seed, tuple(n_frags))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    from instagraal_b200.pyramid_load import pyramid
    out = str(tmp_path / "pyr")
    g = np.load(os.path.join(out, "load.npz"))

    class Rec(RecordedHdf5):
        def __init__(self):
            z = np.load(os.path.join(out, "hdf5_arrays.npz"))
            self.g = {}
            for k in z.files:
                kind, lvl = k.rsplit("_", 1)
                self.g.setdefault(lvl, {})[kind] = z[k] if kind == "data" else np.array([[int(z[k])]], dtype=np.int32)
    pyr = pyramid(out, N_LEVELS, data=Rec())
    for l in range(N_LEVELS):
        lev = pyr.get_level(l)
        assert lev.n_frags == int(g["L%d_n_frags" % l]) and lev.n_contigs == int(g["L%d_n_contigs" % l])
        assert float(lev.mean_value_trans) == float(g["L%d_mean_value_trans" % l]), (l, lev.mean_value_trans)
        assert type(lev.mean_value_trans).__name__ == str(g["L%d_mean_value_trans_type" % l])
        for k in SOA_KEYS:
            assert np.array_equal(lev.S_o_A_frags[k], g["L%d_soa_%s" % (l, k)]), (l, k)
        assert np.array_equal(lev.sparse_mat_csr.data, g["L%d_csr_data" % l]) and np.array_equal(lev.sparse_mat_csr.indices, g["L%d_csr_indices" % l])
        assert np.array_equal(lev.col_vect_frags_4_GL, g["L%d_col_gl" % l]) and np.array_equal(lev.pos_vect_frags_4_GL, g["L%d_pos_gl" % l])
        fd = pyr.spec_level[str(l)]["fragments_dict"]
        for k in FRAG_INFO_KEYS:
            assert [fd[i][k] for i in sorted(fd)] == list(g["L%d_fd_%s" % (l, k)]), (l, k)

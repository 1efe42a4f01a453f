"""TEST INFRASTRUCTURE ONLY: golden vectors for the pyramid LOAD side (PS:1351-1906), made by running the
UNMODIFIED reference classes `pyramid` and `level` of /root/reference/src/instagraal/pyramid_sparse.py on the golden
pyramid folder tests/golden/pyramid/expected/ (itself written by the reference's build functions,
oracle/make_pyramid_golden.py).  h5py is not installed here: `h5py.File` is a read-only stand-in that serves the
(3, nnz) / (1, 1) arrays the reference's fill_sparse_pyramid_level recorded (hdf5_arrays.npz).

   python -m oracle.make_pyramid_load_golden      (here, where /root/reference exists) -> tests/golden/pyramid/load_golden.npz
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FOLDER = os.path.join(ROOT, "tests", "golden", "pyramid", "expected")
OUT = os.path.join(ROOT, "tests", "golden", "pyramid", "load_golden.npz")
N_LEVELS = 4


class NpzBackedFile:
    """what `pyramid` / `level.load_data` read from pyramid.hdf5: data["<lvl>"]["data"] (3, nnz) int32, ["nfrags"] (1, 1)"""

    def __init__(self, path, mode="a"):
        z = np.load(os.path.join(os.path.dirname(path), "hdf5_arrays.npz"))
        self.g = {}
        for k in z.files:
            kind, lvl = k.rsplit("_", 1)
            self.g.setdefault(lvl, {})[kind] = z[k] if kind == "data" else np.array([[int(z[k])]], dtype=np.int32)

    def __getitem__(self, k):
        return self.g[k]

    def __contains__(self, k):
        return k in self.g

    def close(self):
        pass


def reference_module():
    h5 = types.ModuleType("h5py")
    h5.File = NpzBackedFile
    sys.modules["h5py"] = h5
    sys.path.insert(0, os.path.join(HERE, "ref_harness"))   # empty matplotlib stand-in
    sys.path.insert(0, "/root/reference/src")
    import instagraal.pyramid_sparse as PS
    PS.h5py = h5
    return PS


SOA_KEYS = ["pos", "sub_pos", "id_c", "start_bp", "len_bp", "sub_len", "circ", "id", "prev", "next", "l_cont", "sub_l_cont",
            "l_cont_bp", "n_accu"]
FRAG_INFO_KEYS = ["index", "start_pos(bp)", "end_pos(bp)", "size(bp)", "sub_low_index", "sub_high_index", "super_index", "n_accu_frags"]
B_FRAG_INT = ["id_init", "start_pos", "end_pos", "length_kb", "np_id_abs", "curr_id", "contig_id", "init_frag_start", "init_frag_end",
              "sub_frag_start", "sub_frag_end", "super_index", "n_accu_frags"]


def dump(pyr, res):
    """every attribute a caller of the reference reads (simu_single.py, instagraal.py, the sampler constructor)"""
    res["list_contigs_name"] = np.array(pyr.list_contigs_name)
    res["list_contigs_id"] = np.array(pyr.list_contigs_id, dtype=np.int64)
    for l in range(N_LEVELS):
        sp = pyr.spec_level[str(l)]
        fd = sp["fragments_dict"]
        ids = sorted(fd.keys())
        res["L%d_fd_ids" % l] = np.array(ids, dtype=np.int64)
        for k in FRAG_INFO_KEYS:
            res["L%d_fd_%s" % (l, k)] = np.array([fd[i][k] for i in ids], dtype=np.int64)
        res["L%d_fd_init_contig" % l] = np.array([fd[i]["init_contig"] for i in ids])
        res["L%d_fd_tag" % l] = np.array([fd[i]["tag"] for i in ids])
        cd = sp["contigs_dict"]
        res["L%d_cd_keys" % l] = np.array([str(k) for k in cd.keys()])   # insertion order, names popped or not
        frs = [f for k in cd.keys() if isinstance(k, int) for f in cd[k]]
        for k in B_FRAG_INT:
            res["L%d_bf_%s" % (l, k)] = np.array([getattr(f, k) for f in frs], dtype=np.int64)
        res["L%d_bf_gc" % l] = np.array([f.gc_content for f in frs], dtype=np.float64)
        res["L%d_bf_init_name" % l] = np.array([f.init_name for f in frs])
        lev = pyr.get_level(l)
        res["L%d_n_frags" % l] = np.int64(lev.n_frags)
        res["L%d_n_contigs" % l] = np.int64(lev.n_contigs)
        res["L%d_mean_value_trans" % l] = np.float64(lev.mean_value_trans)
        res["L%d_mean_value_trans_type" % l] = np.array(type(lev.mean_value_trans).__name__)
        for k in SOA_KEYS:
            assert lev.S_o_A_frags[k].dtype == np.int32
            res["L%d_soa_%s" % (l, k)] = lev.S_o_A_frags[k]
        res["L%d_vect_frag_np" % l] = np.array([list(t) for t in lev.vect_frag_np.tolist()], dtype=np.int32)
        res["L%d_distri_frag" % l] = lev.distri_frag
        res["L%d_frags_init_contigs" % l] = np.array(lev.frags_init_contigs)
        res["L%d_pos_gl" % l] = lev.pos_vect_frags_4_GL
        res["L%d_col_gl" % l] = lev.col_vect_frags_4_GL
        for nm, m in (("csr", lev.sparse_mat_csr), ("csc", lev.sparse_mat_csc)):
            res["L%d_%s_data" % (l, nm)] = m.data
            res["L%d_%s_indices" % (l, nm)] = m.indices
            res["L%d_%s_indptr" % (l, nm)] = m.indptr
        for c in lev.dict_contigs:
            dc = lev.dict_contigs[c]
            res["L%d_c%d_intra_coord" % (l, c)] = np.array(dc["intra_coord"], dtype=np.int64)
            res["L%d_c%d_tick_kb" % (l, c)] = dc["tick_kb"]
            res["L%d_c%d_end_frags_kb" % (l, c)] = dc["end_frags_kb"]
            res["L%d_c%d_name" % (l, c)] = np.array(dc["name"])
    return res


if __name__ == "__main__":
    PS = reference_module()
    pyr = PS.pyramid(FOLDER, N_LEVELS)
    res = dump(pyr, {})
    np.savez_compressed(OUT, **res)
    for d in (ROOT, os.getcwd()):   # the reference's logger drops a file into the working directory
        for f in os.listdir(d):
            if f.startswith("instagraal-") and f.endswith(".log"):
                os.remove(os.path.join(d, f))
    print("written", OUT, len(res), "arrays;", {l: (int(res["L%d_n_frags" % l]), int(res["L%d_n_contigs" % l]), float(res["L%d_mean_value_trans" % l])) for l in range(N_LEVELS)})

"""Pyramid LOAD side: the classes ``pyramid`` and ``level`` of the reference
(pyramid_sparse.py:1351-1661 and 1663-1961) -- what turns a pyramid folder into the arrays the ``sampler`` constructor is
fed with (``S_o_A_frags``, ``sparse_mat_csr``, ``mean_value_trans``, the fragment dictionaries ``simulation`` reads).

Same class names, constructor arguments, attributes and dictionary keys as the reference, so
``from instagraal_b200.pyramid_load import pyramid`` replaces ``from instagraal.pyramid_sparse import pyramid``.
What is different is how ``level.load_data`` gets there: the reference walks nested Python objects fragment by fragment and
slices the sparse matrix once per contig (row slice, ``tocsc()``, column slice: PS:1866-1886); here the fragment lists are
parsed once into NumPy columns, the 14 scaffold arrays are segment operations on those columns, and the mean trans
contact count is one pass over the stored contacts (contig of the row != contig of the column).

The contact arrays come from ``pyramid.hdf5`` when h5py is importable (group ``"<level>"``: ``data`` (3, nnz) int32,
``nfrags`` (1, 1)); without h5py they are rebuilt from ``level_k/k_abs_frag_contacts.txt`` by
``pyramid_build.fill_sparse_pyramid_level`` (GPU binning, same row order as the reference's HDF5 writer); a caller may also
pass any mapping with that layout as ``data=``.

Reference behaviours kept on purpose (golden vectors dumped from the unmodified reference classes, tests/test_pyramid_load.py):
  * the scaffold arrays are listed contig by contig (order of first appearance of the contig name), not by fragment id;
  * ``sub_l_cont`` is the number of fragments of the same contig one level below (the level itself at level 0);
  * ``mean_value_trans = total_trans / np.float32(n_pairs_trans)`` -- the divisor is rounded to float32 first (PS:1888-1889);
    NaN (a single contig) falls back to ``min(stored values) / 10``;
  * at the top level only contig names that do not parse as an integer are removed from ``contigs_dict`` (PS:1394-1399);
  * ``load_reference_sequence`` drops the LAST line of the FASTA file (``all_lines[start:-1]``, PS:1649) -- the tail of the
    last record is lost exactly as in the reference, so sequences cut from it agree byte for byte.
"""
import colorsys
import gzip
import os

import numpy as np
import scipy.sparse as sp

from . import export as _export


class basic_fragment:
    """fragment.py:4-72 (attribute bag; same attribute names)"""
    __slots__ = ("id_init", "init_contig", "init_name", "start_pos", "end_pos", "length_kb", "gc_content", "np_id_abs",
                 "curr_id", "curr_name", "pos_kb", "contig_id", "orientation", "init_frag_start", "init_frag_end",
                 "sub_frag_start", "sub_frag_end", "super_index", "n_accu_frags")

    @classmethod
    def initiate(cls, np_id_abs, id_init, init_contig, curr_id, start_pos, end_pos, length_kb, gc_content, init_frag_start,
                 init_frag_end, sub_frag_start, sub_frag_end, super_index, id_contig, n_accu_frags):
        o = cls()
        o.id_init, o.init_contig, o.init_name = id_init, init_contig, str(id_init) + "-" + init_contig
        o.start_pos, o.end_pos, o.length_kb, o.gc_content = start_pos, end_pos, length_kb, gc_content
        o.np_id_abs, o.curr_id, o.curr_name, o.pos_kb, o.contig_id, o.orientation = np_id_abs, curr_id, "", 0, id_contig, "w"
        o.init_frag_start, o.init_frag_end, o.sub_frag_start, o.sub_frag_end = init_frag_start, init_frag_end, sub_frag_start, sub_frag_end
        o.super_index, o.n_accu_frags = super_index, n_accu_frags
        return o


class _TextBackedData:
    """stands in for the h5py file when h5py is missing: data["<lvl>"]["data" | "nfrags"] rebuilt from the level's text files"""

    def __init__(self, owner, device):
        self.owner, self.device, self.cache = owner, device, {}

    def __getitem__(self, key):
        if key not in self.cache:
            from .pyramid_build import fill_sparse_pyramid_level
            lvl = int(key)
            n = int(self.owner.spec_level[key]["frag_columns"]["index"].size)
            arr = fill_sparse_pyramid_level(None, lvl, os.path.join(self.owner.spec_level[key]["level_folder"], "%d_abs_frag_contacts.txt" % lvl),
                                            n, device=self.device)
            self.cache[key] = {"data": arr, "nfrags": np.array([[n]], dtype=np.int32)}
        return self.cache[key]

    def __contains__(self, key):
        return key in self.owner.spec_level

    def close(self):
        self.cache.clear()


class pyramid:
    """PS:1351-1661"""

    def __init__(self, pyramid_folder, n_levels, data=None, device=0):
        self.pyramid_folder = pyramid_folder
        self.n_levels = n_levels
        self.pyramid_file = os.path.join(pyramid_folder, "pyramid.hdf5")
        self.spec_level = dict()
        self.struct_initiated = False
        self.resol_F_s_kb = 3
        self.dist_max_kb = 30 * 2 * self.resol_F_s_kb
        for i in range(n_levels):
            level_folder = os.path.join(pyramid_folder, "level_" + str(i))
            sl = self.spec_level[str(i)] = dict()
            sl["level_folder"] = level_folder
            sl["fragments_list_file"] = os.path.join(level_folder, str(i) + "_fragments_list.txt")
            sl["contig_info_file"] = os.path.join(level_folder, str(i) + "_contig_info.txt")
            fd, cd, names, ids, cols = self._read_fragments(sl["fragments_list_file"], i)
            if i == 0:
                self.list_contigs_name, self.list_contigs_id = names, ids
            sl["fragments_dict"], sl["contigs_dict"], sl["frag_columns"] = fd, cd, cols
            if i < n_levels - 1:
                self.update_super_index(fd, os.path.join(level_folder, str(i) + "_sub_2_super_index_frag.txt"))
                self.update_super_index_in_dict_contig(fd, cd)
                cols["super_index"] = np.fromiter((fd[k]["super_index"] for k in range(1, len(fd) + 1)), dtype=np.int64, count=len(fd))
            else:
                for contig_id in list(cd.keys()):
                    try:
                        int(contig_id)
                    except ValueError:
                        cd.pop(contig_id)
        if data is not None:
            self.data = data
        else:
            try:
                import h5py
                self.data = h5py.File(self.pyramid_file, "a")
            except ImportError:
                self.data = _TextBackedData(self, device)

    def close(self):
        self.data.close()

    def get_level(self, level_id):
        return level(self, level_id)

    # -- fragment lists -------------------------------------------------------------------------------------------------
    def _read_fragments(self, fragments_list, lvl):
        """build_frag_dictionnary (PS:1409-1482) + the same table as NumPy columns (file order)"""
        with open(fragments_list, "r") as fh:
            fh.readline()
            rows = [ln.split("\t") for ln in fh if ln]
        n = len(rows)
        contig_dict, list_contigs, list_contigs_id = dict(), [], []
        fragments_info = dict()
        table = np.empty((n, 10), dtype=np.int64)    # index start end size n_accu init_lo init_hi sub_lo sub_hi contig_id
        gc = np.empty(n, dtype=np.float64)
        names_of = [None] * n
        initiate = basic_fragment.initiate
        for k, r in enumerate(rows):
            ci, nm = int(r[0]), r[1]
            start, end, size, g, n_accu, init_lo, init_hi = int(r[2]), int(r[3]), int(r[4]), float(r[5]), int(r[6]), int(r[7]), int(r[8])
            sub_lo, sub_hi = (int(r[9]), int(r[10])) if lvl > 0 else (ci, ci)
            c = contig_dict.get(nm)
            if c is None:
                list_contigs.append(nm)
                list_contigs_id.append(len(list_contigs))
                c = contig_dict[nm] = {"frag": [], "id_contig": len(list_contigs)}
                contig_dict[len(list_contigs)] = []
            cid = c["id_contig"]
            fragments_info[k + 1] = {"init_contig": nm, "index": ci, "tag": r[0] + "-" + nm, "start_pos(bp)": start, "end_pos(bp)": end,
                                     "size(bp)": size, "sub_low_index": sub_lo, "sub_high_index": sub_hi, "super_index": ci,
                                     "n_accu_frags": n_accu}
            f = initiate(k + 1, ci, nm, ci, start, end, size, g, init_lo, init_hi, sub_lo, sub_hi, ci, cid, n_accu)
            c["frag"].append(f)
            contig_dict[cid].append(f)
            table[k] = (ci, start, end, size, n_accu, init_lo, init_hi, sub_lo, sub_hi, cid)
            gc[k] = g
            names_of[k] = nm
        cols = {"index": table[:, 0].copy(), "start": table[:, 1].copy(), "end": table[:, 2].copy(), "size": table[:, 3].copy(), "gc": gc,
                "n_accu": table[:, 4].copy(), "sub_lo": table[:, 7].copy(), "sub_hi": table[:, 8].copy(), "contig_id": table[:, 9].copy(),
                "init_contig": names_of, "super_index": table[:, 0].copy()}
        return fragments_info, contig_dict, list_contigs, list_contigs_id, cols

    def build_frag_dictionnary(self, fragments_list, level):
        return self._read_fragments(fragments_list, level)[:4]

    def update_super_index(self, dict_frag, super_index_file):
        with open(super_index_file, "r") as fh:
            fh.readline()
            for ln in fh:
                if ln:
                    d = ln.split("\t")
                    dict_frag[int(d[0])]["super_index"] = int(d[1])

    def update_super_index_in_dict_contig(self, dict_frag, dict_contig):
        names = set()
        for fr in dict_frag.values():
            names.add(fr["init_contig"])
            dict_contig[dict_contig[fr["init_contig"]]["id_contig"]][fr["index"] - 1].super_index = fr["super_index"]
        for nm in names:
            dict_contig.pop(nm)

    # -- navigation between levels (PS:1512-1628) -----------------------------------------------------------------------
    def zoom_in_frag(self, curr_frag):
        frag, lvl = curr_frag[0], curr_frag[1]
        if lvl <= 0:
            return [curr_frag]
        fd = self.spec_level[str(lvl)]["fragments_dict"][frag]
        return [(i, lvl - 1) for i in range(fd["sub_low_index"], fd["sub_high_index"] + 1)]

    full_zoom_in_frag = zoom_in_frag

    def zoom_out_frag(self, curr_frag):
        frag, lvl = curr_frag[0], curr_frag[1]
        if lvl <= 0:
            return curr_frag
        return (self.spec_level[str(lvl)]["fragments_dict"][frag]["super_index"], lvl + 1)

    def zoom_in_pixel(self, curr_pixel):
        lo, hi, lvl = curr_pixel[0], curr_pixel[1], curr_pixel[2]
        if lvl <= 0:
            return curr_pixel
        fd = self.spec_level[str(lvl)]["fragments_dict"]
        v = [fd[lo]["sub_low_index"], fd[lo]["sub_high_index"], fd[hi]["sub_low_index"], fd[hi]["sub_high_index"]]
        return [min(v), max(v), lvl - 1]

    def zoom_in_area(self, area):
        x, y = area[0], area[1]
        lvl = x[2]
        if not (lvl == y[2] and lvl > 0):
            return area
        hx, hy = self.zoom_in_pixel(x), self.zoom_in_pixel(y)
        return [[min(hx[0], hy[0]), min(hx[1], hy[1]), lvl - 1], [max(hx[0], hy[0]), max(hx[1], hy[1]), lvl - 1]]

    def load_reference_sequence(self, genome_fasta):
        """PS:1630-1660, incl. the dropped last line of the file"""
        opener = gzip.open if str(genome_fasta).endswith(".gz") else open
        with opener(genome_fasta, "rt") as f:
            all_lines = f.readlines()
        self.dict_sequence_contigs = dict()
        heads = [i for i, ln in enumerate(all_lines) if ln[0] == ">" or i == 0]
        for h, nxt in zip(heads, heads[1:] + [len(all_lines) - 1]):
            name = all_lines[h][1:].split()[0].strip()
            body = "".join(all_lines[h + 1:nxt]) if nxt > h else ""
            self.dict_sequence_contigs[name] = body.replace("\n", "").replace("\r", "")


_SOA_KEYS = ("pos", "sub_pos", "id_c", "start_bp", "len_bp", "sub_len", "circ", "id", "prev", "next", "l_cont", "sub_l_cont",
             "l_cont_bp", "n_accu")


class level:
    """PS:1663-2033"""

    def __init__(self, pyramid, level):
        self.level = level
        self.T_frag = np.dtype([(k, np.int32) for k in ("pos", "id_c", "start_bp", "len_bp", "circ", "id", "prev", "next", "l_cont",
                                                        "l_cont_bp", "n_accu")], align=True)
        self.float4 = np.dtype([(k, np.float32) for k in "xyzw"], align=True)
        self.S_o_A_frags = {}
        self.vect_frag_np = []
        self.frags_init_contigs = []
        self.load_data(pyramid)
        self.pyramid = pyramid

    def load_data(self, pyramid):
        grp = pyramid.data[str(self.level)]
        self.n_frags = int(np.copy(grp["nfrags"][0, 0]))
        self.np_2_scipy_sparse = np.copy(grp["data"])
        d = self.np_2_scipy_sparse
        self.sparse_mat_csr = sp.csr_matrix((d[2, :], d[0:2, :]), shape=(self.n_frags, self.n_frags))
        self.sparse_mat_csc = sp.csc_matrix((d[2, :], d[0:2, :]), shape=(self.n_frags, self.n_frags))

        sl = pyramid.spec_level[str(self.level)]
        c = sl["frag_columns"]
        sub = pyramid.spec_level[str(self.level - 1)] if str(self.level - 1) in pyramid.spec_level else sl
        cont_frags, sub_cont_frags = sl["contigs_dict"], sub["contigs_dict"]
        n = c["index"].size
        cid = c["contig_id"]
        n_contigs = len(list(cont_frags.keys()))
        nc = len(pyramid.list_contigs_id)
        # contig by contig, fragments in file order inside a contig (stable)
        order = np.argsort(cid, kind="stable")
        cs = cid[order]
        first = np.r_[0, np.flatnonzero(cs[1:] != cs[:-1]) + 1] if n else np.zeros(0, dtype=np.int64)
        seg_len = np.diff(np.r_[first, n])
        seg_of = np.repeat(np.arange(first.size), seg_len)
        ids = order                                              # np_id_abs - 1
        n_sub = (c["sub_hi"] - c["sub_lo"] + 1)[order]
        size = c["size"][order]
        csum = np.cumsum(n_sub) - n_sub
        sub_pos = csum - csum[first][seg_of] if n else csum
        l_cont_bp = np.add.reduceat(size, first)[seg_of] if n else size
        prev = np.r_[-1, ids[:-1]] if n else ids
        nxt = np.r_[ids[1:], -1] if n else ids
        if n:
            prev[first] = -1
            nxt[np.r_[first[1:] - 1, n - 1]] = -1
        sub_counts = np.array([len(sub_cont_frags[int(k)]) for k in cs[first]], dtype=np.int64)
        S = {"pos": c["index"][order] - 1, "sub_pos": sub_pos, "id_c": cs, "start_bp": c["start"][order], "len_bp": size,
             "sub_len": n_sub, "circ": np.zeros(n, dtype=np.int64), "id": ids, "prev": prev, "next": nxt, "l_cont": seg_len[seg_of],
             "sub_l_cont": sub_counts[seg_of], "l_cont_bp": l_cont_bp, "n_accu": c["n_accu"][order]}
        self.S_o_A_frags = {k: np.array(S[k], dtype=np.int32) for k in _SOA_KEYS}
        v = np.zeros(n, dtype=self.T_frag)
        for k in self.T_frag.names:
            v[k] = self.S_o_A_frags[k]
        self.vect_frag_np = v
        self.distri_frag = np.array(size)
        names = c["init_contig"]
        self.frags_init_contigs = [""] * self.n_frags
        for k in range(n):
            self.frags_init_contigs[k] = names[k]
        # display buffers of the reference's OpenGL viewer (unused by the sampler; kept for attribute compatibility)
        self.pos_vect_frags_4_GL = np.ndarray((np.int32(self.n_frags), 4), dtype=np.float32)
        self.col_vect_frags_4_GL = np.ndarray((np.int32(self.n_frags), 4), dtype=np.float32)
        rgb = np.array([colorsys.hsv_to_rgb(x * 2.5 / n_contigs, 0.5, 0.5) for x in range(n_contigs)], dtype=np.float64).reshape(-1, 3)
        self.pos_vect_frags_4_GL[ids, 0] = (c["index"][order] - 1).astype(np.float32) / np.float32(100.0)
        self.pos_vect_frags_4_GL[ids, 1] = cs.astype(np.float32) / np.float32(100.0)
        self.pos_vect_frags_4_GL[ids, 2] = 0.0
        self.pos_vect_frags_4_GL[ids, 3] = 1.0
        self.col_vect_frags_4_GL[ids, 0:3] = rgb[cs - 1].astype(np.float32)
        self.col_vect_frags_4_GL[ids, 3] = 1.0

        self.dict_contigs = dict()
        tick = c["start"][order] + size / 2.0
        endk = c["end"][order]
        bounds = np.r_[first, n]
        seg_by_id = {int(k): j for j, k in enumerate(cs[first])}
        for id_cont in pyramid.list_contigs_id:
            j = seg_by_id[id_cont]
            a, b = int(bounds[j]), int(bounds[j + 1])
            frs = cont_frags[id_cont]
            self.dict_contigs[id_cont] = {"intra_coord": ids[a:b].tolist(), "frags": frs, "name": frs[0].init_contig,
                                          "tick_kb": np.array(tick[a:b]), "end_frags_kb": np.array(endk[a:b])}

        # mean trans contact count: stored contacts whose two ends lie in different contigs (PS:1863-1891)
        contig_of = np.zeros(self.n_frags, dtype=np.int64)
        contig_of[ids] = cs
        rows, cols_, vals = d[0, :], d[1, :], d[2, :]
        total_trans = np.int64(vals[contig_of[rows] != contig_of[cols_]].sum(dtype=np.int64))
        n_tot_intra = 0
        for ln in seg_len[:nc].tolist():
            n_tot_intra += ln * (ln - 1) / 2
        n_tot = self.n_frags * (self.n_frags - 1) / 2 - n_tot_intra
        with np.errstate(divide="ignore", invalid="ignore"):
            self.mean_value_trans = total_trans / np.float32(n_tot)
        if np.isnan(self.mean_value_trans):
            self.mean_value_trans = np.amin(self.sparse_mat_csr.data) / 10.0
        self.n_contigs = len(self.dict_contigs)

    def build_seq_per_bin(self, genome_fasta):
        """PS:1938-1961"""
        self.pyramid.load_reference_sequence(genome_fasta)
        cont_frags = self.pyramid.spec_level[str(self.level)]["contigs_dict"]
        self.list_seq = []
        for cont in sorted(k for k in cont_frags.keys() if not isinstance(k, str)):
            for frag in cont_frags[cont]:
                self.list_seq.append(self.pyramid.dict_sequence_contigs[frag.init_contig][frag.start_pos:frag.end_pos])

    def generate_new_fasta(self, vect_frags, new_fasta, info_frags):
        """PS:1963-2033 -> export.generate_new_fasta (byte-compatible, tests/test_export.py)"""
        _export.generate_new_fasta(self, vect_frags, new_fasta, info_frags)

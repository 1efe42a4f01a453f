"""Pyramid build: the x3 binning of one pyramid level into the next, with the reference's
file formats byte for byte.

Mirrors ``instagraal/pyramid_sparse.py`` (same function names, same arguments, same files written):

  * ``init_frag_list``            PS:399-465   level-0 fragment list in the pyramid's 9-column layout
  * ``subsample_data_set``        PS:468-724   contig list, fragment list, contact file and old->new index of the next level
  * ``fill_sparse_pyramid_level`` PS:331-397   the (3, nnz) int32 array + fragment count of a level's HDF5 group
  * ``build``                     PS:178-277   the level loop (text levels always; ``pyramid.hdf5`` when h5py is importable)
  * ``remove_problematic_fragments`` PS:731-1030  the filtering step (locked fragments merged forward or destroyed)
  * ``build_and_filter``          PS:30-176    what ``simulation.select_data_set`` calls: unfiltered level, filter, level loop, load

The reference walks every contact line through nested Python dictionaries, once per level (hours at 1e8 contacts).  Here the
bookkeeping of contigs / fragments is vectorised NumPy on the host and the contacts are binned on the GPU by
``ig_bin_contacts`` (stable radix sort of the ordered pair keys + segmented sums, include/instagraal_b200.h): there is no
CPU path for them, the call fails loudly without a CUDA device.  Reference quirk Q13 is reproduced: ``subsample_data_set``
reads the header with ``readline()`` and then starts its loop at index 1 of ``readlines()``, i.e. the FIRST DATA LINE of every
contact file is dropped (PS:679-683).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil

import numpy as np

from . import _lib as L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def file_len(fname):
    """PS:23-27"""
    i = -1
    with open(fname) as f:
        for i, _l in enumerate(f):
            pass
    return i + 1


def bin_contacts(fa, fb, nc, old2new=None, first_appearance_order=False, device=0):
    """(a, b, n) of the binned contact list: ids mapped through ``old2new`` (0-based -> 0-based), pair ordered, equal pairs
    summed; sorted by (a, b), or rows ascending with the columns in order of first appearance (the HDF5 layout)."""
    fa = np.ascontiguousarray(fa, dtype=np.int32)
    fb = np.ascontiguousarray(fb, dtype=np.int32)
    nc = np.ascontiguousarray(nc, dtype=np.int32)
    n = len(fa)
    out_a = np.empty(max(n, 1), dtype=np.int32)
    out_b = np.empty(max(n, 1), dtype=np.int32)
    out_n = np.empty(max(n, 1), dtype=np.int64)
    n_out = C.c_int64(0)
    m = np.ascontiguousarray(old2new, dtype=np.int32) if old2new is not None else None
    rc = L.lib().ig_bin_contacts(int(device), n, _ptr(fa), _ptr(fb), _ptr(nc), _ptr(m) if m is not None else None,
                                 len(m) if m is not None else 0, int(bool(first_appearance_order)),
                                 _ptr(out_a), _ptr(out_b), _ptr(out_n), C.byref(n_out))
    L.check(None, rc, "ig_bin_contacts")
    k = n_out.value
    return out_a[:k], out_b[:k], out_n[:k]


def _read_contacts(path, skip_first_data_line):
    """the three integer columns of a contact file (header skipped; Q13: optionally also the first data line)"""
    import pandas as pd
    df = pd.read_csv(path, sep="\t", header=0, names=["a", "b", "n"], dtype=np.int64, engine="c",
                     skiprows=[1] if skip_first_data_line else None)
    return df["a"].to_numpy(), df["b"].to_numpy(), df["n"].to_numpy()


def _write_contacts(path, a, b, n):
    import pandas as pd
    with open(path, "w") as h:
        h.write("%s\t%s\t%s\n" % ("id_frag_a", "id_frag_b", "n_contact"))
        if len(a):
            pd.DataFrame({"a": a, "b": b, "n": n}).to_csv(h, sep="\t", header=False, index=False, lineterminator="\n")


def init_frag_list(fragment_list, new_frag_list):
    """PS:399-465: returns the number of fragments"""
    with open(fragment_list, "r") as hin, open(new_frag_list, "w") as hout:
        hout.write("%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\n" % ("id", "chrom", "start_pos", "end_pos", "size", "gc_content",
                                                            "accu_frag", "frag_start", "frag_end"))
        hin.readline()
        i = 0
        for line_frag in hin:
            i += 1
            data = line_frag.split("\t")
            hout.write("%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\n" % (data[0], data[1], data[2], data[3], data[4], str(float(data[5])),
                                                                str(1), data[0], data[0]))
    return i


def subsample_data_set(contig_info, fragments_list, fact_sub_sample, abs_fragments_contacts, new_abs_fragments_contacts_file,
                       min_bin_per_contig, new_contig_list_file, new_fragments_list_file, old_2_new_file, device=0):
    """PS:468-724, same arguments and outputs (``device`` = CUDA ordinal of the contact binning)."""
    if fact_sub_sample <= 1:
        shutil.copy(fragments_list, new_fragments_list_file)
        shutil.copy(contig_info, new_contig_list_file)
        shutil.copy(abs_fragments_contacts, new_abs_fragments_contacts_file)
        nfrags = file_len(fragments_list) - 1
        with open(old_2_new_file, "w") as h:
            h.write("%s\t%s\n" % ("current_id", "super_id"))
            for ind in range(0, nfrags):
                h.write("%s\t%s\n" % (str(ind + 1), str(ind + 1)))
        return nfrags
    fact = int(fact_sub_sample)
    # ---- contigs: which are binned, new relative / absolute ids (PS:498-566)
    names, lengths, n_frags_c = [], [], []
    with open(contig_info, "r") as h:
        h.readline()
        for line in h:
            data = line.split("\t")
            names.append(data[0]); lengths.append(data[1]); n_frags_c.append(int(data[2]))
    n_frags_c = np.asarray(n_frags_c, dtype=np.int64)
    binned = (n_frags_c.astype(np.float32) / np.float32(fact)) >= min_bin_per_contig   # PS:520: python int / np.float32 -> float32
    n_new_c = np.where(binned, (n_frags_c + fact - 1) // fact, n_frags_c)
    cum_new = np.concatenate([[0], np.cumsum(n_new_c)])
    cum_old = np.concatenate([[0], np.cumsum(n_frags_c)])
    with open(new_contig_list_file, "w") as h:
        h.write("%s\t%s\t%s\t%s\n" % ("contig", "length_kb", "n_frags", "cumul_length"))
        for c in range(len(names)):
            h.write("%s\t%s\t%s\t%s\n" % (names[c], lengths[c], int(n_new_c[c]), int(cum_new[c])))
    n_old, n_new = int(cum_old[-1]), int(cum_new[-1])
    contig_of_old = np.repeat(np.arange(len(names)), n_frags_c)
    rel_old = np.arange(n_old) - cum_old[contig_of_old]                      # 0-based position inside the contig
    new_rel0 = np.where(binned[contig_of_old], rel_old // fact, rel_old)
    old2new0 = (cum_new[contig_of_old] + new_rel0).astype(np.int64)          # 0-based new absolute id per old fragment
    first_old = np.flatnonzero(np.concatenate([[True], old2new0[1:] != old2new0[:-1]])) if n_old else np.zeros(0, np.int64)
    last_old = np.concatenate([first_old[1:] - 1, [n_old - 1]]) if n_old else np.zeros(0, np.int64)
    # ---- fragments (PS:572-667)
    start_pos = np.empty(n_old, dtype=np.int64); end_pos = np.empty(n_old, dtype=np.int64)
    gc = np.empty(n_old, dtype=np.float64)
    ifs = np.empty(n_old, dtype=np.int64); ife = np.empty(n_old, dtype=np.int64)
    with open(fragments_list, "r") as h:
        h.readline()
        i = 0
        for line in h:
            data = line.split("\t")
            start_pos[i] = int(data[2]); end_pos[i] = int(data[3]); gc[i] = float(data[5])
            ifs[i] = int(data[7]); ife[i] = int(data[8])
            i += 1
    if i != n_old:
        raise ValueError("fragments list has %d records, contig info says %d" % (i, n_old))
    counts = last_old - first_old + 1
    # np.array(list).mean() (PS:631): NumPy adds fewer than 8 numbers strictly left to right (np.add.reduceat does NOT: it
    # differs in the last bit), longer runs pairwise -- those go through np.mean itself
    if fact < 8:
        gc_sum = gc[first_old].copy() if n_old else gc
        for j in range(1, fact):
            idx = np.minimum(first_old + j, max(n_old - 1, 0))
            gc_sum = gc_sum + np.where(j < counts, gc[idx], 0.0)
        gc_mean = gc_sum / counts
    else:
        gc_mean = np.array([np.array(gc[a:b + 1].tolist()).mean() for a, b in zip(first_old, last_old)])
    new_contig = contig_of_old[first_old]
    new_rel = (np.arange(n_new) - cum_new[new_contig]) + 1
    with open(new_fragments_list_file, "w") as h:
        h.write("%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\n" % ("id", "chrom", "start_pos", "end_pos", "size", "gc_content", "accu_frag",
                                                                  "init_frag_start", "init_frag_end", "sub_frag_start", "sub_frag_end"))
        for k in range(n_new):
            f0, f1 = int(first_old[k]), int(last_old[k])
            sp, ep = int(start_pos[f0]), int(end_pos[f1])
            h.write("%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\n" % (
                str(int(new_rel[k])), names[int(new_contig[k])], sp, ep, ep - sp, np.float64(gc_mean[k]),
                str(int(ife[f1]) - int(ifs[f0]) + 1), int(ifs[f0]), int(ife[f1]), f0 + 1, f1 + 1))
    # ---- contacts (PS:670-722), binned on the GPU
    if not (abs_fragments_contacts == "SIMU"):
        fa, fb, nc = _read_contacts(abs_fragments_contacts, skip_first_data_line=True)   # quirk Q13
        a, b, n = bin_contacts(fa, fb, nc, old2new=old2new0, first_appearance_order=False, device=device)
        _write_contacts(new_abs_fragments_contacts_file, a, b, n)
    with open(old_2_new_file, "w") as h:
        h.write("%s\t%s\n" % ("current_id", "super_id"))
        h.write("".join("%d\t%d\n" % (i + 1, old2new0[i] + 1) for i in range(n_old)))
    return n_new


def fill_sparse_pyramid_level(pyramid_handle, level, contact_file, nfrags, device=0):
    """PS:331-397.  Returns the (3, nnz) int32 array; when ``pyramid_handle`` is an h5py file (or anything with its
    create_group / create_dataset interface) the group ``str(level)`` is written exactly like the reference does."""
    fa, fb, nc = _read_contacts(contact_file, skip_first_data_line=False)
    a, b, n = bin_contacts(fa, fb, nc, old2new=None, first_appearance_order=True, device=device)
    arr = np.zeros((3, len(a)), dtype=np.int32)
    arr[0, :] = a; arr[1, :] = b; arr[2, :] = n
    if pyramid_handle is not None:
        level_hdf5 = pyramid_handle.create_group(str(level))
        data_2_sparse = level_hdf5.create_dataset("data", (3, arr.shape[1]), "i")
        data_nfrags = level_hdf5.create_dataset("nfrags", (1, 1), "i")
        data_2_sparse[0, :] = arr[0]
        data_2_sparse[1, :] = arr[1]
        data_2_sparse[2, :] = arr[2]
        data_nfrags[:] = nfrags
    return arr


def build(base_folder, size_pyramid, factor, min_bin_per_contig, output_folder=None, device=0):
    """PS:178-277: same folder layout, same text files per level.  ``pyramid.hdf5`` is written when h5py is importable
    (the reference requires it); without h5py the text levels are still built and the HDF5 cache is skipped."""
    try:
        import h5py
    except ImportError:
        h5py = None
    pyramid_root = output_folder if output_folder is not None else base_folder
    contig_info = os.path.join(base_folder, "info_contigs.txt")
    fragments_list = os.path.join(base_folder, "fragments_list.txt")
    init_abs_fragments_contacts = os.path.join(base_folder, "abs_fragments_contacts_weighted.txt")
    all_pyramid_folder = os.path.join(pyramid_root, "pyramids")
    pyramid_folder = os.path.join(all_pyramid_folder, "pyramid_" + str(size_pyramid) + "_no_thresh")
    os.makedirs(pyramid_folder, exist_ok=True)
    pyramid_handle = h5py.File(os.path.join(pyramid_folder, "pyramid.hdf5"), "a") if h5py is not None else None
    level_folder = os.path.join(pyramid_folder, "level_0")
    os.makedirs(level_folder, exist_ok=True)
    current_contig_info = os.path.join(level_folder, "0_contig_info.txt")
    current_frag_list = os.path.join(level_folder, "0_fragments_list.txt")
    current_abs_fragments_contacts = os.path.join(level_folder, "0_abs_frag_contacts.txt")
    sub_2_super = None
    nfrags_per_level = []
    for level in range(0, size_pyramid):
        level_folder = os.path.join(pyramid_folder, "level_" + str(level))
        os.makedirs(level_folder, exist_ok=True)
        pre = str(level) + "_"
        if level == 0:
            shutil.copyfile(contig_info, current_contig_info)
            shutil.copyfile(init_abs_fragments_contacts, current_abs_fragments_contacts)
            nfrags = init_frag_list(fragments_list, current_frag_list)
            new_contacts, new_contigs, new_frags = current_abs_fragments_contacts, current_contig_info, current_frag_list
        else:
            new_contigs = os.path.join(level_folder, pre + "contig_info.txt")
            new_frags = os.path.join(level_folder, pre + "fragments_list.txt")
            new_contacts = os.path.join(level_folder, pre + "abs_frag_contacts.txt")
            if all(os.path.exists(p) for p in (new_contigs, new_frags, new_contacts, sub_2_super)):
                nfrags = file_len(new_frags) - 1
            else:
                nfrags = subsample_data_set(current_contig_info, current_frag_list, factor, current_abs_fragments_contacts, new_contacts,
                                            min_bin_per_contig, new_contigs, new_frags, sub_2_super, device=device)
        if pyramid_handle is not None:
            try:
                status = pyramid_handle.attrs[str(level)] == "done"
            except KeyError:
                pyramid_handle.attrs[str(level)] = "pending"
                status = False
            if not status:
                fill_sparse_pyramid_level(pyramid_handle, level, new_contacts, nfrags, device=device)
                pyramid_handle.attrs[str(level)] = "done"
        nfrags_per_level.append(nfrags)
        current_frag_list, current_contig_info, current_abs_fragments_contacts = new_frags, new_contigs, new_contacts
        sub_2_super = os.path.join(level_folder, pre + "sub_2_super_index_frag.txt")
    if pyramid_handle is not None:
        pyramid_handle.close()
    return nfrags_per_level


def remove_problematic_fragments(contig_info, fragments_list, abs_fragments_contacts, new_contig_list_file, new_fragments_list_file,
                                 new_abs_fragments_contacts_file, pyramid, thresh_factor=1, device=0):
    """PS:731-1030, same arguments, files and return value (the sparsity threshold): the filtering step every real run starts
    with (``build_and_filter``).  Fragments whose row of the symmetrised level-0 matrix is too empty (<= mean - thresh_factor *
    std of the fill ratio), implausibly full (> mean + 50 std) or shorter than 50 bp are *locked*: a locked fragment is merged
    into the next unlocked fragment of its contig, a trailing run of locked fragments is destroyed together with its contacts.

    ``pyramid`` is the open HDF5 file of the unfiltered 1-level pyramid or any mapping with ``["0"]["data"]`` (3, nnz) and
    ``["0"]["nfrags"]`` (1, 1).  The reference's two diagnostic PDF plots (written into the working directory) are not made.

    Reference behaviours kept on purpose: a contig starts where the index column says 1; the first fragment written for a
    contig starts at 0 whatever the file says, later ones at the END of the previously written fragment as spelled in the file;
    ``accu_frag`` is the only accumulator that is not reset at a contig start, so the counts of a destroyed trailing run leak
    into the first fragment written for the next contig (PS:880-893 vs. 948); the ``size <= 1`` lock never takes effect
    (PS:909-910 sets the flag that the write overwrites); mean GC content via ``np.array(list).mean()``."""
    import scipy.sparse as sp
    lvl = pyramid["0"]
    d = np.asarray(lvl["data"])
    nfrags = int(np.asarray(lvl["nfrags"])[0, 0])
    m = sp.csr_matrix((d[2, :], d[0:2, :]), shape=(nfrags, nfrags))
    full = m + m.transpose()
    sparsity = np.float32(np.diff(full.indptr)) / np.float32(nfrags)
    mean_s, std_s = sparsity.mean(), sparsity.std()
    thresh_max = mean_s + 50 * std_s
    thresh = mean_s - thresh_factor * std_s
    # ---- the fragment table (file order)
    with open(fragments_list, "r") as h:
        h.readline()
        rows = [ln.split("\t") for ln in h if ln]
    n_old = len(rows)
    idx = np.fromiter((int(r[0]) for r in rows), dtype=np.int64, count=n_old)
    chrom = [r[1] for r in rows]
    end_s = [r[3] for r in rows]
    size = np.fromiter((int(r[4]) for r in rows), dtype=np.int64, count=n_old)
    gc = np.fromiter((float(r[5]) for r in rows), dtype=np.float64, count=n_old)
    accu = np.fromiter((int(r[6]) for r in rows), dtype=np.int64, count=n_old)
    length = np.fromiter((int(r[3]) - int(r[2]) for r in rows), dtype=np.int64, count=n_old)
    lock = np.zeros(n_old, dtype=bool)
    lock[np.flatnonzero(sparsity <= thresh)] = True
    lock[np.flatnonzero(sparsity > thresh_max)] = True
    lock[length < 50] = True
    # ---- who is written, who merges into whom, who is destroyed
    seg = np.cumsum(idx == 1)                         # contig segment of every line
    wr = ~lock
    widx = np.flatnonzero(wr)                         # lines that close a new fragment
    n_new = widx.size
    grp = np.cumsum(wr) - wr                          # new fragment (0-based) a line would end up in
    has_writer = grp < n_new
    writer_seg = np.full(n_old, -1, dtype=np.int64)
    writer_seg[has_writer] = seg[widx[grp[has_writer]]]
    merged = has_writer & (writer_seg == seg)         # same contig segment as the closing line
    destroyed = ~merged
    new_size = np.bincount(grp[merged], weights=size[merged], minlength=n_new).astype(np.int64)[:n_new]
    new_accu = np.bincount(grp[has_writer], weights=accu[has_writer], minlength=n_new).astype(np.int64)[:n_new]   # (not reset at a contig start)
    n_run = np.bincount(grp[merged], minlength=n_new)[:n_new]
    first_in_seg = np.ones(n_new, dtype=bool)
    if n_new > 1:
        first_in_seg[1:] = seg[widx[1:]] != seg[widx[:-1]]
    seg_start = np.flatnonzero(first_in_seg)
    new_rel = np.arange(n_new) - np.repeat(seg_start, np.diff(np.r_[seg_start, n_new])) + 1
    merged_idx = np.flatnonzero(merged)
    run_start = np.searchsorted(grp[merged_idx], np.arange(n_new), side="left")
    names, n_new_c, len_c = [], {}, {}
    with open(contig_info, "r") as h:
        h.readline()
        for ln in h:
            if ln:
                dat = ln.split("\t")
                int(dat[1]); int(dat[2]); int(dat[3])     # the reference parses them (get_contig_info_from_file)
                names.append(dat[0]); n_new_c[dat[0]] = 0; len_c[dat[0]] = 0
    with open(new_fragments_list_file, "w") as h:
        h.write("%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\n" % ("id", "chrom", "start_pos", "end_pos", "size", "gc_content", "accu_frag",
                                                            "frag_start", "frag_end"))
        out = []
        for k in range(n_new):
            i = int(widx[k])
            if n_run[k] == 1:
                g = np.float64(gc[i])
            else:
                g = np.array(gc[merged_idx[run_start[k]:run_start[k] + n_run[k]]].tolist()).mean()
            start = 0 if first_in_seg[k] else end_s[int(widx[k - 1])]
            out.append("%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\t%s\n" % (str(int(new_rel[k])), chrom[i], start, end_s[i], str(int(new_size[k])), str(g),
                                                                int(new_accu[k]), str(int(new_rel[k])), str(int(new_rel[k]))))
            n_new_c[chrom[i]] += 1
            len_c[chrom[i]] += int(new_size[k])
        h.write("".join(out))
    with open(new_contig_list_file, "w") as h:
        h.write("%s\t%s\t%s\t%s\n" % ("contig", "length_kb", "n_frags", "cumul_length"))
        cumul = 0
        for c in names:
            if n_new_c[c] > 0:
                h.write("%s\t%s\t%s\t%s\n" % (c, str(len_c[c]), str(n_new_c[c]), str(cumul)))
                cumul += n_new_c[c]
    # ---- contacts: both ends must survive; re-binned on the GPU
    fa, fb, nc = _read_contacts(abs_fragments_contacts, skip_first_data_line=False)
    keep = ~(destroyed[fa] | destroyed[fb])
    a, b, n = bin_contacts(fa[keep], fb[keep], nc[keep], old2new=np.where(destroyed, 0, grp), first_appearance_order=False, device=device)
    _write_contacts(new_abs_fragments_contacts_file, a, b, n)
    return thresh


def build_and_filter(base_folder, size_pyramid, factor, thresh_factor=1, output_folder=None, device=0):
    """PS:30-176: the unfiltered 1-level pyramid, the filtered level 0 (``remove_problematic_fragments``), then the level loop,
    in the folder layout of the reference (``pyramids/pyramid_1_no_thresh``, ``pyramids/pyramid_<n>_thresh_auto``); returns the
    loaded ``pyramid`` object (``pyramid_load.pyramid``).  ``pyramid.hdf5`` is written and read when h5py is importable;
    without it the text levels are the cache and the contact arrays are rebuilt from them on load."""
    from .pyramid_load import pyramid
    try:
        import h5py
    except ImportError:
        h5py = None
    min_bin_per_contig = 1
    pyramid_root = output_folder if output_folder is not None else base_folder
    all_pyramid_folder = os.path.join(pyramid_root, "pyramids")
    os.makedirs(all_pyramid_folder, exist_ok=True)
    init_pyramid_folder = os.path.join(all_pyramid_folder, "pyramid_1_no_thresh")
    if not os.path.exists(init_pyramid_folder):
        build(base_folder, 1, factor, min_bin_per_contig, output_folder=pyramid_root, device=device)
    init0 = os.path.join(init_pyramid_folder, "level_0")
    contig_info = os.path.join(init0, "0_contig_info.txt")
    fragments_list = os.path.join(init0, "0_fragments_list.txt")
    init_contacts = os.path.join(init0, "0_abs_frag_contacts.txt")
    pyramid_folder = os.path.join(all_pyramid_folder, "pyramid_" + str(size_pyramid) + "_thresh_auto")
    level_folder = os.path.join(pyramid_folder, "level_0")
    os.makedirs(level_folder, exist_ok=True)
    cur_contigs = os.path.join(level_folder, "0_contig_info.txt")
    cur_frags = os.path.join(level_folder, "0_fragments_list.txt")
    cur_contacts = os.path.join(level_folder, "0_abs_frag_contacts.txt")
    if not all(os.path.exists(p) for p in (cur_contigs, cur_frags, cur_contacts)):
        if h5py is not None:
            pyramid_0 = h5py.File(os.path.join(init_pyramid_folder, "pyramid.hdf5"), "r")
        else:
            n0 = file_len(fragments_list) - 1
            pyramid_0 = {"0": {"data": fill_sparse_pyramid_level(None, 0, init_contacts, n0, device=device),
                               "nfrags": np.array([[n0]], dtype=np.int32)}}
        remove_problematic_fragments(contig_info, fragments_list, init_contacts, cur_contigs, cur_frags, cur_contacts, pyramid_0,
                                     thresh_factor=thresh_factor, device=device)
        if h5py is not None:
            pyramid_0.close()
    pyramid_handle = h5py.File(os.path.join(pyramid_folder, "pyramid.hdf5"), "a") if h5py is not None else None
    sub_2_super = os.path.join(level_folder, "0_sub_2_super_index_frag.txt")
    for level in range(0, size_pyramid):
        level_folder = os.path.join(pyramid_folder, "level_" + str(level))
        os.makedirs(level_folder, exist_ok=True)
        pre = str(level) + "_"
        new_contigs = os.path.join(level_folder, pre + "contig_info.txt")
        new_frags = os.path.join(level_folder, pre + "fragments_list.txt")
        new_contacts = os.path.join(level_folder, pre + "abs_frag_contacts.txt")
        if level > 0 and not all(os.path.exists(p) for p in (new_contigs, new_frags, new_contacts, sub_2_super)):
            nfrags = subsample_data_set(cur_contigs, cur_frags, factor, cur_contacts, new_contacts, min_bin_per_contig, new_contigs,
                                        new_frags, sub_2_super, device=device)
        else:
            nfrags = file_len(new_frags) - 1
        if pyramid_handle is not None:
            try:
                status = pyramid_handle.attrs[str(level)] == "done"
            except KeyError:
                pyramid_handle.attrs[str(level)] = "pending"
                status = False
            if not status:
                fill_sparse_pyramid_level(pyramid_handle, level, new_contacts, nfrags, device=device)
                pyramid_handle.attrs[str(level)] = "done"
        cur_frags, cur_contigs, cur_contacts = new_frags, new_contigs, new_contacts
        sub_2_super = os.path.join(level_folder, pre + "sub_2_super_index_frag.txt")
    if pyramid_handle is not None:
        pyramid_handle.close()
    return pyramid(pyramid_folder, size_pyramid, device=device)

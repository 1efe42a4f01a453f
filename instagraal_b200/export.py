"""Per-cycle output of the scaffold: ``genome.fasta`` + ``info_frags.txt``.

Drop-in for ``level.generate_new_fasta(vect_frags, new_fasta, info_frags)`` of the reference
(pyramid_sparse.py:1963-2033), byte-compatible output, written for large assemblies: one global sort
instead of a ``np.nonzero`` scan per contig, sequences joined once instead of grown by ``+=`` per fragment,
fragment coordinates cached as arrays instead of looked up in nested dicts per fragment per cycle.

``level`` is duck-typed like the reference's ``pyramid_sparse.level``:
    level.level                               pyramid level (int)
    level.frags_init_contigs[i]               name of the initial contig of fragment i
    level.pyramid.spec_level[str(level.level)]["fragments_dict"][i + 1]["start_pos(bp)" | "end_pos(bp)"]
    level.pyramid.dict_sequence_contigs[name] sequence (anything sliceable to str)
``vect_frags`` needs the arrays id_c, pos, ori, activ, id_d (sampler.gpu_vect_frags after copy_from_gpu()).

Reference behaviours kept on purpose: contigs listed in ascending id in info_frags.txt and by decreasing
length (stable) in the FASTA; only contigs whose fragments are all active are written; reverse strands are
complemented with the table TAGCtagc -> ATCGATCG (lower case becomes upper case, other letters unchanged);
lines of 61 characters; and the last line is dropped when it would hold exactly one character
(``idx_cut_EOL[-1] != len_seq - 1``, pyramid_sparse.py:2025).
"""
import numpy as np

_COMPLEMENT = str.maketrans("TAGCtagc", "ATCGATCG")
_LINE = 61


def _frag_tables(level):
    cache = getattr(level, "_ig_b200_frag_tables", None)
    if cache is None:
        fd = level.pyramid.spec_level[str(level.level)]["fragments_dict"]
        n = len(level.frags_init_contigs)
        start = np.fromiter((fd[i + 1]["start_pos(bp)"] for i in range(n)), dtype=np.int64, count=n)
        end = np.fromiter((fd[i + 1]["end_pos(bp)"] for i in range(n)), dtype=np.int64, count=n)
        cache = (start, end, list(level.frags_init_contigs))
        try:
            level._ig_b200_frag_tables = cache
        except AttributeError:
            pass
    return cache


def generate_new_fasta(level, vect_frags, new_fasta, info_frags):
    id_c = np.asarray(vect_frags.id_c)
    pos = np.asarray(vect_frags.pos)
    ori = np.asarray(vect_frags.ori)
    activ = np.asarray(vect_frags.activ)
    id_d = np.asarray(vect_frags.id_d)
    start_bp, end_bp, init_contigs = _frag_tables(level)
    seqs = level.pyramid.dict_sequence_contigs

    order = np.lexsort((pos, id_c))                      # by contig id, then position
    ids, first = np.unique(id_c[order], return_index=True)
    bounds = np.r_[first, order.size]
    all_active = np.minimum.reduceat((activ[order] == 1).astype(np.int8), first) == 1

    contig_seq = {}
    ok = []
    with open(info_frags, "w") as h_info:
        for ci, cid in enumerate(ids):
            if not all_active[ci]:
                continue
            ok.append(cid)
            fr = order[bounds[ci]:bounds[ci + 1]]
            init = id_d[fr]
            names = [init_contigs[j] for j in init]
            st = start_bp[init].tolist()
            en = end_bp[init].tolist()
            oris = ori[fr].tolist()
            lines = [">3C-assembly|contig_%s\n" % cid, "init_contig\tid_frag\torientation\tstart\tend\n"]
            lines += ["%s\t%s\t%s\t%s\t%s\n" % (nm, j, o, s, e) for nm, j, o, s, e in zip(names, init.tolist(), oris, st, en)]
            h_info.write("".join(lines))
            parts = []
            for nm, o, s, e in zip(names, oris, st, en):
                piece = str(seqs[nm][s:e])
                parts.append(piece[::-1].translate(_COMPLEMENT) if o == -1 else piece)
            contig_seq[cid] = "".join(parts)

    with open(new_fasta, "w") as h_fa:
        for cid in sorted(ok, key=lambda c: len(contig_seq[c]), reverse=True):
            seq = contig_seq[cid]
            n = len(seq)
            out = [">3C-assembly-contig_%s\n" % cid]
            if n > 0:
                last = ((n - 1) // _LINE) * _LINE        # start of the last (possibly partial) line
                out += [seq[i:i + _LINE] + "\n" for i in range(0, last, _LINE)]
                if last != n - 1:                         # reference quirk: a 1-character last line is dropped
                    out.append(seq[last:] + "\n")
            h_fa.write("".join(out))

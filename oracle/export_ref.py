"""TEST INFRASTRUCTURE ONLY -- literal restatement of the reference's per-cycle FASTA / info_frags writer,
``level.generate_new_fasta`` (pyramid_sparse.py:1963-2033), statement by statement (per-contig np.nonzero scan,
argsort of the positions, string growth by +=, the 61-character line splitter with its last-line test).
The reference module itself cannot be imported here (needs h5py/Biopython), so parity of
instagraal_b200/export.py is anchored on this restatement: PARITY UNPINNED by reference outputs."""
import numpy as np


def generate_new_fasta(level, vect_frags, new_fasta, info_frags):
    id_c_frag = vect_frags.id_c
    pos_frag = vect_frags.pos
    ori_frag = vect_frags.ori
    activ_frag = vect_frags.activ
    handle_new_fasta = open(new_fasta, "w")
    handle_info_frags = open(info_frags, "w")
    list_id_contigs = np.unique(id_c_frag)                       # PS:1973
    list_seq_new_contigs = dict()
    list_contigs_ok = []
    fd = level.pyramid.spec_level[str(level.level)]["fragments_dict"]
    for id_cont in list_id_contigs:                              # PS:1977-2012
        list_frags = np.nonzero(id_c_frag == id_cont)[0]
        if np.all(activ_frag[list_frags] == 1):
            list_contigs_ok.append(id_cont)
            header = ">3C-assembly|contig_" + str(id_cont)
            handle_info_frags.write("%s\n" % (header))
            handle_info_frags.write("%s\t%s\t%s\t%s\t%s\n" % ("init_contig", "id_frag", "orientation", "start", "end"))
            new_positions = pos_frag[list_frags]
            ordered_id_frags = list_frags[np.argsort(new_positions)]
            list_seq_new_contigs[id_cont] = ""
            for f in ordered_id_frags:
                ori = ori_frag[f]
                init_frag_id = vect_frags.id_d[f]
                init_contig = level.frags_init_contigs[init_frag_id]
                start_bp = fd[init_frag_id + 1]["start_pos(bp)"]
                end_bp = fd[init_frag_id + 1]["end_pos(bp)"]
                extract_seq = level.pyramid.dict_sequence_contigs[init_contig][start_bp:end_bp]
                if ori == -1:
                    seq = extract_seq[::-1]
                    seq = seq.translate(str.maketrans("TAGCtagc", "ATCGATCG"))
                else:
                    seq = extract_seq
                handle_info_frags.write("%s\t%s\t%s\t%s\t%s\n" % (init_contig, str(init_frag_id), str(ori), str(start_bp), str(end_bp)))
                list_seq_new_contigs[id_cont] += seq

    def contig_length(c):
        return len(list_seq_new_contigs[c])

    for id_cont in sorted(list_contigs_ok, key=contig_length, reverse=True):   # PS:2014-2028
        cont_seq = list_seq_new_contigs[id_cont]
        header = ">3C-assembly-contig_" + str(id_cont)
        handle_new_fasta.write("%s\n" % (header))
        len_line = 61
        len_seq = len(cont_seq)
        if len_seq > 0:
            idx_cut_EOL = list(range(0, len_seq, len_line))
            for id_s in range(1, len(idx_cut_EOL)):
                line = cont_seq[idx_cut_EOL[id_s - 1]: idx_cut_EOL[id_s]]
                handle_new_fasta.write("%s\n" % (line))
            if idx_cut_EOL[-1] != len_seq - 1:
                line = cont_seq[idx_cut_EOL[-1]:]
                handle_new_fasta.write("%s\n" % (line))
    handle_new_fasta.close()
    handle_info_frags.close()

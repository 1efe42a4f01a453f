"""TEST INFRASTRUCTURE -- builds tests/golden/yeast_toy_geometry.npz, the fragment geometry of BASELINE.json
configs[0] ("tests/data toy assembly ... level 4"), from the reference's own test data.

Runs only where /root/reference exists (this container); the fixture it writes is committed and travels.

What it restates (no reference code is executed or copied):
  * instagraal-pre's digestion (pre.py -> hicstuff digest): cut sites of DpnII (^GATC) and HinfI (G^ANTC) found by
    regex on every contig of tests/data/yeast.contigs.fa.gz -> level-0 restriction fragments;
  * the pyramid's x3 binning (pyramid_sparse.py:523-549, min_bin_per_contig = 1, pyramid_sparse.py:54): per contig,
    consecutive groups of 3 fragments (the last group keeps the remainder); contigs with fewer than 3 fragments
    are copied unbinned.  Applied 4 times -> level 4 fragments whose sub-fragments are the level-3 fragments;
  * the true layout recorded in the contig descriptions "from_<chrom>:<start>-<end>"
    (scripts/make_insilico_assembly.py:72-79).
Contacts are not part of the fixture: instagraal_b200.synth simulates them on this geometry (seeded).
"""
import gzip
import os
import re
import sys

import numpy as np

REF_FASTA = "/root/reference/tests/data/yeast.contigs.fa.gz"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "yeast_toy_geometry.npz")


def read_fasta(path):
    name, desc, seq = None, None, []
    with gzip.open(path, "rt") as fh:
        for line in fh:
            if line.startswith(">"):
                if name is not None:
                    yield name, desc, "".join(seq)
                parts = line[1:].strip().split(None, 1)
                name, desc, seq = parts[0], (parts[1] if len(parts) > 1 else ""), []
            else:
                seq.append(line.strip().upper())
    if name is not None:
        yield name, desc, "".join(seq)


def digest(seq):
    cuts = {0, len(seq)}
    for m in re.finditer("GATC", seq):          # DpnII  ^GATC
        cuts.add(m.start())
    for m in re.finditer("(?=G[ACGT]?.TC)", seq):  # HinfI  G^ANTC (lookahead: overlapping sites)
        if re.match("GA[ACGTN]TC", seq[m.start():m.start() + 5]):
            cuts.add(m.start() + 1)
    c = np.array(sorted(cuts), dtype=np.int64)
    return np.diff(c)[np.diff(c) > 0]


def bin3(groups):
    """groups: list (one per contig) of arrays of fragment lengths -> (binned lengths, children per bin)"""
    out_len, out_n = [], []
    for g in groups:
        n = len(g)
        if n >= 3:
            st = np.arange(0, n, 3)
            out_len.append(np.add.reduceat(g, st))
            out_n.append(np.diff(np.r_[st, n]))
        else:
            out_len.append(g.copy())
            out_n.append(np.ones(n, dtype=np.int64))
    return out_len, out_n


def main():
    if not os.path.exists(REF_FASTA):
        sys.exit("reference test data not available here")
    names, chroms, starts, level = [], [], [], []
    for name, desc, seq in read_fasta(REF_FASTA):
        m = re.search(r"from_([^:]+):(\d+)-(\d+)", desc)
        names.append(name); chroms.append(m.group(1)); starts.append(int(m.group(2)))
        level.append(digest(seq))
    n0 = sum(len(g) for g in level)
    for _ in range(3):
        level, _ = bin3(level)
    sub = level                                   # level 3 = sub-fragments of level 4
    frag_len, frag_nsub = bin3(sub)                # level 4
    chrom_names = sorted(set(chroms))
    chrom_idx = np.array([chrom_names.index(c) for c in chroms], dtype=np.int32)
    np.savez_compressed(
        OUT,
        contig_names=np.array(names), chrom_names=np.array(chrom_names), contig_chrom=chrom_idx,
        contig_start=np.array(starts, dtype=np.int64),
        frags_per_contig=np.array([len(g) for g in frag_len], dtype=np.int32),
        frag_nsub=np.concatenate(frag_nsub).astype(np.int32),
        sub_len_bp=np.concatenate(sub).astype(np.int64),
        n_level0=np.int64(n0),
    )
    print("contigs", len(names), "level-0 fragments", n0, "level-3", sum(len(g) for g in sub), "level-4",
          sum(len(g) for g in frag_len), "->", os.path.normpath(OUT))


if __name__ == "__main__":
    main()

"""N4: the per-cycle FASTA / info_frags writer is byte-identical to the restated reference
(oracle/export_ref.py follows pyramid_sparse.py:1963-2033 statement by statement)."""
import filecmp
import time
import types

import numpy as np
import pytest

from instagraal_b200 import export
from oracle import export_ref
from oracle.fuzz import random_state


def _fake_level(rng, n_frags, n_init_contigs, frag_len):
    # initial contigs made of consecutive fragments; lengths chosen so that contig lengths hit every residue mod 61
    per = np.sort(rng.choice(n_init_contigs, n_frags))
    names = ["ctg_%03d" % c for c in per]
    lens = frag_len(n_frags)
    fd, seqs, cursor = {}, {}, {}
    for i in range(n_frags):
        s = cursor.get(names[i], 0)
        fd[i + 1] = {"start_pos(bp)": int(s), "end_pos(bp)": int(s + lens[i])}
        cursor[names[i]] = s + lens[i]
    for nm, ln in cursor.items():
        seqs[nm] = "".join(rng.choice(list("ACGTacgtNn"), int(ln), p=[.22, .22, .22, .22, .02, .02, .02, .02, .02, .02]))
    lvl = types.SimpleNamespace(level=4, frags_init_contigs=names,
                                pyramid=types.SimpleNamespace(spec_level={"4": {"fragments_dict": fd}}, dict_sequence_contigs=seqs))
    return lvl


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_fasta_and_info_frags_byte_identical(tmp_path, seed):
    rng = np.random.RandomState(seed)
    n = 300
    lvl = _fake_level(rng, n, 12, lambda k: rng.randint(1, 140, k))
    st = random_state(n, rng, p_circ=0.2)
    vf = types.SimpleNamespace(id_c=st["id_c"], pos=st["pos"], ori=st["ori"], id_d=rng.permutation(n).astype(np.int32),
                               activ=(rng.rand(n) > 0.01).astype(np.int32))
    export_ref.generate_new_fasta(lvl, vf, str(tmp_path / "ref.fa"), str(tmp_path / "ref.txt"))
    export.generate_new_fasta(lvl, vf, str(tmp_path / "new.fa"), str(tmp_path / "new.txt"))
    assert filecmp.cmp(tmp_path / "ref.fa", tmp_path / "new.fa", shallow=False)
    assert filecmp.cmp(tmp_path / "ref.txt", tmp_path / "new.txt", shallow=False)
    assert (tmp_path / "new.fa").stat().st_size > 1000


def test_line_splitter_quirk_every_residue(tmp_path):
    """single-fragment contigs of every length 0..130: lengths with len % 61 == 1 lose their last character
    in the reference (pyramid_sparse.py:2025) -- kept."""
    rng = np.random.RandomState(5)
    n = 131
    lvl = _fake_level(rng, n, n, lambda k: np.arange(k))
    lvl.frags_init_contigs = ["ctg_%03d" % i for i in range(n)]
    fd = lvl.pyramid.spec_level["4"]["fragments_dict"]
    for i in range(n):
        fd[i + 1] = {"start_pos(bp)": 0, "end_pos(bp)": i}
        lvl.pyramid.dict_sequence_contigs["ctg_%03d" % i] = "".join(rng.choice(list("ACGT"), i))
    vf = types.SimpleNamespace(id_c=np.arange(n, dtype=np.int32), pos=np.zeros(n, np.int32), ori=np.where(np.arange(n) % 2, -1, 1).astype(np.int32),
                               id_d=np.arange(n, dtype=np.int32), activ=np.ones(n, np.int32))
    export_ref.generate_new_fasta(lvl, vf, str(tmp_path / "ref.fa"), str(tmp_path / "ref.txt"))
    export.generate_new_fasta(lvl, vf, str(tmp_path / "new.fa"), str(tmp_path / "new.txt"))
    assert filecmp.cmp(tmp_path / "ref.fa", tmp_path / "new.fa", shallow=False)
    assert filecmp.cmp(tmp_path / "ref.txt", tmp_path / "new.txt", shallow=False)


def test_export_is_faster_on_a_large_scaffold(tmp_path):
    """20,000 fragments in 40 contigs (~60 Mb): same bytes, and the writer does not scan per contig."""
    rng = np.random.RandomState(3)
    n = 20000
    lvl = _fake_level(rng, n, 200, lambda k: rng.randint(1500, 4500, k))
    id_c = np.sort(rng.randint(0, 40, n)).astype(np.int32)
    pos = np.zeros(n, np.int32)
    for c in np.unique(id_c):
        m = np.flatnonzero(id_c == c)
        pos[m] = rng.permutation(m.size)
    vf = types.SimpleNamespace(id_c=id_c, pos=pos, ori=rng.choice([-1, 1], n).astype(np.int32), id_d=rng.permutation(n).astype(np.int32),
                               activ=np.ones(n, np.int32))
    t0 = time.perf_counter()
    export_ref.generate_new_fasta(lvl, vf, str(tmp_path / "ref.fa"), str(tmp_path / "ref.txt"))
    t_ref = time.perf_counter() - t0
    t0 = time.perf_counter()
    export.generate_new_fasta(lvl, vf, str(tmp_path / "new.fa"), str(tmp_path / "new.txt"))
    t_new = time.perf_counter() - t0
    assert filecmp.cmp(tmp_path / "ref.fa", tmp_path / "new.fa", shallow=False)
    assert filecmp.cmp(tmp_path / "ref.txt", tmp_path / "new.txt", shallow=False)
    print("generate_new_fasta: restated reference %.2f s, instagraal_b200.export %.2f s" % (t_ref, t_new))


@pytest.mark.parametrize("seed", [0, 1])
def test_against_files_written_by_the_reference_method(tmp_path, seed):
    """pinned: tests/golden/export/* were written by the reference's OWN level.generate_new_fasta (pyramid_sparse.py:1963-2033,
    called unbound on a stand-in level by oracle/make_export_golden.py); the product and the restatement reproduce them byte
    for byte from the same seeded inputs."""
    import os
    from oracle.make_export_golden import make_case, stand_in_level
    names, starts, ends, seqs, vf = make_case(seed)
    lvl = stand_in_level(names, starts, ends, seqs)
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "export")
    for impl, tag in ((export, "new"), (export_ref, "ref")):
        fa, tx = str(tmp_path / (tag + ".fa")), str(tmp_path / (tag + ".txt"))
        impl.generate_new_fasta(lvl, types.SimpleNamespace(**vf), fa, tx)
        assert filecmp.cmp(fa, os.path.join(g, "genome_%d.fasta" % seed), shallow=False), tag
        assert filecmp.cmp(tx, os.path.join(g, "info_frags_%d.txt" % seed), shallow=False), tag


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference/src/instagraal"), reason="differential run against the live reference (build container only)")
def test_differential_against_the_live_reference_method(tmp_path):
    """more seeds and shapes than the two committed goldens, written by the reference's own method in a child process:
    one contig, many one-fragment contigs, a larger scaffold"""
    import os
    import pickle
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cases = [(7, 1, 1), (8, 60, 60), (9, 500, 23), (10, 37, 3), (11, 900, 2)]
    code = (
        "import sys, os, types, pickle\n"
        "sys.path.insert(0, %r)\n"
        "from oracle.make_export_golden import make_case, stand_in_level\n"
        "from oracle.make_pyramid_golden import reference_module\n"
        "PS = reference_module()\n"
        "for seed, n, n_init in %r:\n"
        "    names, starts, ends, seqs, vf = make_case(seed, n=n, n_init=n_init)\n"
        "    PS.level.generate_new_fasta(stand_in_level(names, starts, ends, seqs), types.SimpleNamespace(**vf), 'g_%%d.fa' %% seed, 'i_%%d.txt' %% seed)\n"
    ) % (root, cases)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    from oracle.make_export_golden import make_case, stand_in_level
    for seed, n, n_init in cases:
        names, starts, ends, seqs, vf = make_case(seed, n=n, n_init=n_init)
        fa, tx = str(tmp_path / ("mine_%d.fa" % seed)), str(tmp_path / ("mine_%d.txt" % seed))
        export.generate_new_fasta(stand_in_level(names, starts, ends, seqs), types.SimpleNamespace(**vf), fa, tx)
        assert filecmp.cmp(fa, str(tmp_path / ("g_%d.fa" % seed)), shallow=False), seed
        assert filecmp.cmp(tx, str(tmp_path / ("i_%d.txt" % seed)), shallow=False), seed

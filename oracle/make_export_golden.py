"""TEST INFRASTRUCTURE ONLY: golden genome.fasta / info_frags.txt for N4, written by the reference's OWN
``level.generate_new_fasta`` (pyramid_sparse.py:1963-2033, called unbound on a stand-in ``self`` that carries exactly the
attributes the method reads) for seeded random scaffolds.   python -m oracle.make_export_golden  -> tests/golden/export/"""
import os
import types

import numpy as np

from oracle.fuzz import random_state
from oracle.make_pyramid_golden import reference_module

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "export")


def make_case(seed, n=240, n_init=9):
    rng = np.random.RandomState(seed)
    per = np.sort(rng.choice(n_init, n))
    names = ["ctg_%03d" % c for c in per]
    lens = rng.randint(1, 140, n)
    starts, ends, cursor = np.zeros(n, np.int64), np.zeros(n, np.int64), {}
    for i in range(n):
        s = cursor.get(names[i], 0)
        starts[i], ends[i] = s, s + lens[i]
        cursor[names[i]] = s + lens[i]
    seqs = {nm: "".join(rng.choice(list("ACGTacgtNn"), int(ln), p=[.22, .22, .22, .22, .02, .02, .02, .02, .02, .02])) for nm, ln in cursor.items()}
    st = random_state(n, rng, p_circ=0.2)
    vf = dict(id_c=st["id_c"], pos=st["pos"], ori=st["ori"], id_d=rng.permutation(n).astype(np.int32),
              activ=(rng.rand(n) > 0.01).astype(np.int32))
    return names, starts, ends, seqs, vf


def stand_in_level(names, starts, ends, seqs):
    fd = {i + 1: {"start_pos(bp)": int(starts[i]), "end_pos(bp)": int(ends[i])} for i in range(len(names))}
    return types.SimpleNamespace(level=4, frags_init_contigs=list(names),
                                 pyramid=types.SimpleNamespace(spec_level={"4": {"fragments_dict": fd}}, dict_sequence_contigs=seqs))


if __name__ == "__main__":
    PS = reference_module()
    os.makedirs(OUT, exist_ok=True)
    for seed in (0, 1):
        names, starts, ends, seqs, vf = make_case(seed)
        lvl = stand_in_level(names, starts, ends, seqs)
        PS.level.generate_new_fasta(lvl, types.SimpleNamespace(**vf), os.path.join(OUT, "genome_%d.fasta" % seed),
                                    os.path.join(OUT, "info_frags_%d.txt" % seed))
        print("written case", seed)
    root = os.path.dirname(HERE)
    for f in os.listdir(root):
        if f.startswith("instagraal-") and f.endswith(".log"):
            os.remove(os.path.join(root, f))

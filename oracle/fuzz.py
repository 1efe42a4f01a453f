"""TEST INFRASTRUCTURE ONLY: random (but invariant-respecting) scaffolds, including
circular contigs, for fuzzing move semantics against the reference kernels."""
from __future__ import annotations

import numpy as np


def random_state(n, rng, p_circ=0.3, max_contigs=None):
    """Returns a live-state dict (13 int32 fields) with contig ids 0..NC-1."""
    max_contigs = max_contigs or max(1, n // 3)
    nc = int(rng.randint(1, max_contigs + 1))
    assign = np.sort(rng.randint(0, nc, n))
    _, assign = np.unique(assign, return_inverse=True)
    nc = assign.max() + 1
    perm = rng.permutation(n)
    s = {k: np.zeros(n, dtype=np.int32) for k in ("pos", "sub_pos", "id_c", "start_bp", "len_bp", "sub_len",
                                                    "circ", "prev", "next", "l_cont", "sub_l_cont", "l_cont_bp", "ori")}
    s["len_bp"][:] = rng.randint(500, 20000, n)
    s["sub_len"][:] = rng.randint(1, 4, n)
    s["ori"][:] = rng.choice([-1, 1], n)
    label = rng.permutation(nc)
    for c in range(nc):
        members = perm[assign == c]
        L = len(members)
        circ = int(L >= 3 and rng.rand() < p_circ)
        bp = sp = 0
        for i, f in enumerate(members):
            s["pos"][f] = i
            s["sub_pos"][f] = sp
            s["start_bp"][f] = bp
            s["id_c"][f] = label[c]
            s["circ"][f] = circ
            s["prev"][f] = members[i - 1] if i > 0 else (members[-1] if circ else -1)
            s["next"][f] = members[i + 1] if i < L - 1 else (members[0] if circ else -1)
            bp += s["len_bp"][f]
            sp += s["sub_len"][f]
        s["l_cont"][members] = L
        s["sub_l_cont"][members] = sp
        s["l_cont_bp"][members] = bp
    return s

"""CPU fuzz of the integer move semantics:
  (1) oracle restatement (oracle/moves.py) vs the reference's OWN kernels compiled for the CPU
      (oracle/_ref/libref_cpu.so), on random scaffolds incl. circular contigs;
  (2) the product's on-the-fly per-fragment functions (instagraal_b200/csrc/ig_moves.cuh, host
      build) vs the oracle."""
import ctypes
import os

import numpy as np
import pytest

from oracle import moves as mv
from oracle import ref_kernels as rk
from oracle.fuzz import random_state

HOST_LIB = os.path.join(os.path.dirname(__file__), "host_build", "libig_moves_host.so")


@pytest.mark.skipif(not rk.available(), reason="oracle/_ref/libref_cpu.so not built")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_moves_vs_reference_kernels(seed):
    rng = np.random.RandomState(seed)
    for _ in range(60):
        n = int(rng.randint(2, 40))
        st = random_state(n, rng)
        max_id = int(st["id_c"].max())
        a, b = [int(x) for x in rng.choice(n, 2, replace=False)]
        ref, valid_ref, stale = rk.perform_mutations(st, a, b, max_id)
        assert not stale, "reference left struct entries unwritten (quirk Q4 reached)"
        mine, valid = mv.perform_mutations(st, a, b, max_id)
        assert valid == valid_ref
        for m in range(24):
            for k in mv.FIELDS:
                assert np.array_equal(mine[m][k], ref[m][k]), (seed, a, b, m, k)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_product_move_functions_vs_oracle(built, seed):
    lib = ctypes.CDLL(HOST_LIB)
    rng = np.random.RandomState(100 + seed)
    for _ in range(100):
        n = int(rng.randint(2, 48))
        st = random_state(n, rng)
        max_id = int(st["id_c"].max())
        a, b = [int(x) for x in rng.choice(n, 2, replace=False)]
        prev_valid = rng.choice([-1, 1, 0], 12).astype(np.int32)
        fe = int(rng.randint(0, 2))
        inp = np.ascontiguousarray(np.stack([st[k] for k in mv.FIELDS]).astype(np.int32))
        out = np.zeros((24, 13, n), dtype=np.int32)
        valid = np.zeros(12, np.int32)
        uniq = np.zeros(24, np.int32)
        q4 = ctypes.c_int(0)
        vp = ctypes.c_void_p
        nu = lib.ig_host_eval_all(n, inp.ctypes.data_as(vp), a, b, max_id, prev_valid.ctypes.data_as(vp), fe,
                                  out.ctypes.data_as(vp), valid.ctypes.data_as(vp), uniq.ctypes.data_as(vp),
                                  ctypes.byref(q4))
        ref, valid_ref = mv.perform_mutations(st, a, b, max_id)
        assert list(uniq[:nu]) == mv.extract_uniq_mutations(st, a, b, prev_valid.tolist(), fe)
        assert valid.tolist() == valid_ref
        for m in range(24):
            for i, k in enumerate(mv.FIELDS):
                assert np.array_equal(out[m, i], ref[m][k]), (seed, a, b, m, k)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_rigid_motion_classes(built, seed):
    """Every op moves the fragments between two breakpoints rigidly: the motion of each fragment of the
    affected contigs equals the motion of its class representative (what the scoring kernel's
    class-pair table is built from), for linear and circular contigs, all 24 ops."""
    lib = ctypes.CDLL(HOST_LIB)
    rng = np.random.RandomState(500 + seed)
    vp = ctypes.c_void_p
    total = 0
    for it in range(300):
        n = int(rng.randint(2, 96 if it % 3 else 400))
        st = random_state(n, rng)
        max_id = int(st["id_c"].max())
        a, b = [int(x) for x in rng.choice(n, 2, replace=False)]
        inp = np.ascontiguousarray(np.stack([st[k] for k in mv.FIELDS]).astype(np.int32))
        ncls, nchk = ctypes.c_int(0), ctypes.c_int(0)
        bad = lib.ig_host_check_classes(n, inp.ctypes.data_as(vp), a, b, max_id, ctypes.byref(ncls), ctypes.byref(nchk))
        assert bad == 0, (seed, it, n, a, b)
        total += nchk.value
    assert total > 10000


def test_scaffold_invariants_after_moves():
    """SURVEY A.3 invariants hold for every op applied to a valid scaffold (linear contigs)."""
    rng = np.random.RandomState(7)
    for _ in range(40):
        n = int(rng.randint(3, 30))
        st = random_state(n, rng, p_circ=0.0)
        a, b = [int(x) for x in rng.choice(n, 2, replace=False)]
        muts, _ = mv.perform_mutations(st, a, b, int(st["id_c"].max()))
        for m, s in enumerate(muts):
            assert (s["pos"] >= 0).all() and (s["l_cont"] > 0).all(), m
            assert (s["l_cont_bp"] > s["start_bp"]).all() and (s["start_bp"] >= 0).all(), m
            assert ((s["start_bp"] == 0) == (s["pos"] == 0)).all(), m
            for c in np.unique(s["id_c"]):
                mem = np.flatnonzero(s["id_c"] == c)
                assert sorted(s["pos"][mem].tolist()) == list(range(len(mem))), (m, c)
                assert (s["l_cont"][mem] == len(mem)).all(), (m, c)

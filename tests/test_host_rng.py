"""The facade's neighbour draw is stream-identical to the reference's np.random.choice call
(cuda_lib_gl_single.py:3117-3124): same values, same generator state afterwards."""
import numpy as np

from instagraal_b200.host_rng import prepare, weighted_choice_no_replace
from instagraal_b200.synth import WORKLOADS, make_level


def _state():
    st = np.random.get_state()
    return st[1].copy(), st[2]


def test_weighted_choice_matches_numpy_draw_for_draw():
    rng = np.random.RandomState(0)
    for trial in range(4000):
        n = int(rng.randint(1, 70))
        xk = rng.permutation(2000)[:n].astype(np.int32)
        dat = rng.poisson(3, n).astype(np.float32) * np.float32(3.0)
        if rng.rand() < 0.3:
            dat[rng.randint(n)] = 900.0      # one dominant weight: repeated draws, NumPy's loop continues
        if dat.sum() == 0:
            dat[:] = 1
        pk = dat / np.linalg.norm(dat, 1)
        size = min(int(rng.randint(1, 9)), int(np.count_nonzero(pk)))
        p64, cdf0 = prepare(pk)
        seed = int(rng.randint(1 << 30))
        np.random.seed(seed)
        want = np.random.choice(xk, size, p=pk, replace=False)
        sw = _state()
        np.random.seed(seed)
        got = weighted_choice_no_replace(xk, p64, cdf0, size)
        sg = _state()
        assert np.array_equal(want, got), trial
        assert sw[1] == sg[1] and np.array_equal(sw[0], sg[0]), trial


def test_on_a_level_distribution():
    """the (xk, pk) of a synthetic level, as setup_distri_frags builds them (cuda_lib_gl_single.py:3053-3101)"""
    level = make_level(WORKLOADS["toy"])
    sym = (level.sub_sampled_sparse_matrix + level.sub_sampled_sparse_matrix.T).tocsr()
    np.random.seed(3)
    for i in range(level.n_frags):
        st, en = sym.indptr[i], sym.indptr[i + 1]
        vk, yk = sym.data[st:en], sym.indices[st:en]
        het = np.nonzero(yk != i)[0]
        xk = np.copy(yk)[het]
        if len(xk) == 0:
            continue
        dat = np.float32(np.copy(vk)[het]) * 3.0
        pk = dat / np.linalg.norm(dat, 1) if dat.sum() > 0 else np.ones_like(dat) / len(dat)
        n = min(5, int(np.count_nonzero(pk)))
        state = np.random.get_state()
        want = np.random.choice(xk, n, p=pk, replace=False)
        after = _state()
        np.random.set_state(state)
        got = weighted_choice_no_replace(xk, *prepare(pk), n)
        assert np.array_equal(want, got) and _state()[1] == after[1]

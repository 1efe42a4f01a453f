"""B-ref route (i): the UNMODIFIED reference ``sampler`` class on the GPU (oracle/ref_unmodified.py: a verbatim copy of the
reference package + a stand-in for pycuda over cuda-python + the reference kernels' cubin) in lockstep with the product.
Teacher = the reference class: every step starts from its state, with its candidates.  Scores agree to 1e-9 relative except
on the uniq positions >= n_sub % 64, whose value in the reference depends on the run-to-run order of slice_sp_mat's atomics
(see tests/test_bench_configs.py); chosen move and the 13 x NF state are compared whenever the choice is the same."""
import numpy as np
import pytest

from conftest import load_golden
from instagraal_b200.synth import WORKLOADS, make_level
from parity_common import state_mismatch_modulo_length_ties

pytestmark = pytest.mark.gpu


def test_unmodified_reference_class_on_gpu_vs_product(built):
    from oracle import ref_unmodified as ru
    if not ru.available():
        pytest.skip("baseline/_ref copy of the reference or its cubin not built")
    from test_gpu_parity import GpuImpl
    g = load_golden("toy_bomb_seed2")
    level = make_level(WORKLOADS["toy"])
    p8 = g["params8"]
    ref = ru.UnmodifiedSampler(level, p8)
    mine = GpuImpl(level)
    mine.set_params(p8)
    ref.set_state(g["state0"])
    state = np.ascontiguousarray(g["state0"], dtype=np.int32)
    n_same = n_order = 0
    n_steps = 60
    for t in range(n_steps):
        a = int(g["step_A"][t])
        cands = [int(c) for c in g["step_cands"][t][:int(g["step_ncand"][t])]]
        valid = ref.s.gpu_list_valid_insert.get().copy()
        mine.set_state(state)
        mine.set_valid(valid)
        o, dist, op, b, mean_len, nc = ref.step_sampler(a, cands)
        sa = ref.all_scores
        r = mine.step(a, cands)
        sb = np.asarray(r["scores"], dtype=np.float64)
        assert np.array_equal(sa != 0, sb != 0), (t, a, cands)
        nz = sa != 0
        rel = np.zeros_like(sa)
        rel[nz] = np.abs(sa[nz] - sb[nz]) / np.abs(sa[nz])
        bad = np.nonzero(rel >= 1e-9)[0]
        for gidx in bad:   # only where the reference's own value is order-dependent
            k, u = divmod(int(gidx), 24)
            pos_in_uniq = int(np.count_nonzero(nz[k * 24:k * 24 + u]))
            rmod = int(mine.s.n_sub_vals[k]) % 64
            assert rmod > 0 and pos_in_uniq >= rmod, (t, k, u, pos_in_uniq, rmod, float(rel[gidx]))
        new_state = ref.get_state()
        if len(bad):
            n_order += 1
        elif (int(op), int(b)) == (int(r["op"]), int(r["B"])):
            n_same += 1
            assert not state_mismatch_modulo_length_ties(mine.get_state(), new_state), (t, op, b)
            assert float(r["dist"]) == float(dist) and int(r["n_contigs"]) == int(nc)
        state = new_state
    assert n_same >= 0.7 * n_steps, (n_same, n_order)
    mine.s.free_gpu()

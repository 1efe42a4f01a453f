"""The facade driven exactly like the reference driver loop (instagraal.py:196-289, the loop oracle/make_golden.py
recorded the golden vectors with): same seed, same call order => the facade must consume NumPy's legacy global
stream draw for draw like the unmodified reference class did.  Checked against tests/golden/*.npz:

  * bomb_the_genome()          -> the bombed scaffold equals golden `state0` bit for bit (a17),
  * the visiting order and every step's candidate list equal the golden ones (host RNG order, a3),
  * step_nuisance_parameters() -> all 7 returned fields, the proposed test parameters and the likelihood under
    them equal golden `step_nuis` / `step_params` (a18): RNG order (choice(4), normal, rand), fsolve, accept.

The chain state is teacher-forced per step (set_state to the golden pre-step state; exact ties may break
differently, see parity_common), the RNG stream is NOT touched: its consumption does not depend on the chain state.
"""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden
from instagraal_b200.synth import WORKLOADS, make_level

pytestmark = pytest.mark.gpu

CASES = {  # oracle/make_golden.py CASES: (n_cycles, nuis_after)
    "micro_seed0": (3, 70),
    "micro_bomb_seed1": (4, 110),
    "toy_bomb_seed2": (2, 200),
}


@pytest.mark.parametrize("overlap", [True, False])
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_facade_replays_reference_driver_loop(built, name, overlap):
    """overlap = the facade prepares the nuisance proposal (RNG draws + fsolve) on a worker thread while the GPU scores
    the step before it (the default) / strictly in place: same stream, same values either way."""
    from test_gpu_parity import make_sampler
    g = load_golden(name)
    level = make_level(WORKLOADS[str(g["workload"])])
    n_cycles, nuis_after = CASES[name]
    np.random.seed(int(g["seed"]))
    s = make_sampler(level)
    s.overlap_nuisance_proposal = overlap
    max_kb, bin_kb, _ = g["hist_args"]
    s.estimate_parameters_rippe(max_kb, bin_kb, False)
    assert np.allclose(np.array(list(s.param_simu[0]), dtype=np.float32), g["params8"], rtol=1e-5)
    s.set_param_simu(g["params8"])   # the fit agrees to 1e-5 (scipy leastsq); continue from the reference's exact values
    s.param_simu_test = s.param_simu
    if int(g["bomb"]):
        s.bomb_the_genome()
    if int(g["bomb"]):
        assert np.array_equal(s._get_state(), g["state0"]), "bombed scaffold differs from the reference's"
    else:   # the reference's initial contig labels are the level's own (1-based, not yet renumbered): same partition
        st0 = s._get_state()
        assert all(np.array_equal(st0[i], g["state0"][i]) for i in range(13) if i != 2)
        assert len(np.unique(np.stack([st0[2], g["state0"][2]]), axis=1)[0]) == len(np.unique(st0[2]))
    n_steps = len(g["step_A"])
    nuis_at = {int(t): k for k, t in enumerate(g["step_nuis_step"])}
    list_frags = np.arange(0, s.n_new_frags)
    dt = np.float32(0.01)
    prev_state = g["state0"]
    t = 0
    n_nuis = 0
    for j in range(n_cycles):
        s.gpu_vect_frags.copy_from_gpu()
        np.random.shuffle(list_frags)
        for id_frag in list_frags:
            if t >= n_steps:
                break
            assert int(id_frag) == int(g["step_A"][t]), ("visiting order", t)
            s._set_state(prev_state)
            s.set_valid_insert(g["step_valid_before"][t])
            got_p = np.array(list(s.param_simu[0]), dtype=np.float32)
            assert np.array_equal(got_p, g["step_params_before"][t]), ("parameters before step", t, got_p, g["step_params_before"][t])
            s.step_sampler(int(id_frag), 5, dt)
            nc = int(g["step_ncand"][t])
            assert list(s.candidates) == [int(c) for c in g["step_cands"][t][:nc]], ("candidate draw", t)
            if t >= nuis_after:
                k = nuis_at[t]
                s.likelihood_t = np.float64(g["step_o"][t])   # teacher-forced (a tie may have been broken differently)
                fact, d, d_max, d_nuc, slope, lik, success, _y = s.step_nuisance_parameters(dt, t, n_cycles * int(s.n_new_frags))
                want = g["step_nuis"][k]
                assert np.array_equal(np.array(list(s.param_simu_test[0]), dtype=np.float32), g["step_params"][k]), ("test parameters", t)
                got_lik_nuis = float(np.ravel(s.likelihood_nuis)[0])
                assert abs(got_lik_nuis - want[7]) <= 2e-8 * abs(want[7]) + 1e-9, ("nuisance likelihood", t, got_lik_nuis, want[7])
                assert int(success) == int(want[6]), ("accept / reject", t)
                for got_v, want_v, nm in ((fact, want[0], "fact"), (d, want[1], "d"), (d_max, want[2], "d_max"),
                                          (d_nuc, want[3], "d_nuc"), (slope, want[4], "slope")):
                    assert np.float32(got_v) == np.float32(want_v), (nm, t, got_v, want_v)
                assert abs(float(np.ravel(lik)[0]) - want[5]) <= 2e-8 * abs(want[5]) + 1e-9, ("likelihood_t", t)
                n_nuis += 1
            prev_state = g["step_states"][t]
            t += 1
    assert t == n_steps and n_nuis == len(g["step_nuis_step"])
    assert s.n_nuis_overlapped == (max(n_nuis - 1, 0) if overlap else 0)   # every nuisance step but the first found its proposal ready
    # the RNG stream position after the whole run: one more draw must equal the reference's next draw, which the
    # golden file does not hold -- instead the visiting order / candidates / nuisance draws above pin every consumed value
    s.free_gpu()

"""TEST INFRASTRUCTURE ONLY: generates tests/golden/*.npz by running the UNMODIFIED reference
``sampler`` class (imported from /root/reference/src) on the CPU emulation of its own kernels
(oracle/ref_harness + oracle/_ref/libref_cpu.so, built by oracle/Makefile).

Run here (the container that has /root/reference):   python -m oracle.make_golden
The driver loop mirrors instagraal.py:196-289 (full_em): bomb, per-cycle shuffle, step_sampler per
fragment, step_nuisance_parameters after each step once enabled.  The host RNG is NumPy's legacy
global stream seeded right before the run, as SURVEY F.3 prescribes.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
FIELDS13 = ("pos", "sub_pos", "id_c", "start_bp", "len_bp", "sub_len", "circ", "prev", "next",
            "l_cont", "sub_l_cont", "l_cont_bp", "ori")


def _import_reference():
    sys.path.insert(0, os.path.join(HERE, "ref_harness"))
    sys.path.insert(0, "/root/reference/src")
    sys.path.insert(0, ROOT)
    import instagraal.cuda_lib_gl_single as CL  # noqa: E402  (the reference, unmodified)
    return CL


def snapshot(gs):
    gs.copy_from_gpu()
    return np.stack([np.array(getattr(gs, k), dtype=np.int32) for k in FIELDS13])


def run_reference(level, seed, n_cycles, bomb, max_steps, nuis_after, n_struct_dumps=6, state_every=1):
    CL = _import_reference()
    np.random.seed(seed)
    s = CL.sampler(*level.sampler_args())
    id_start = np.nonzero(s.gpu_vect_frags.start_bp == 0)[0]
    max_dist_kb = s.gpu_vect_frags.l_cont_bp[id_start].max() / 1000.0
    mean_size_bin_kb = level.S_o_A_sub_frags["len_bp"].mean() / 1000.0
    s.estimate_parameters_rippe(max_dist_kb, mean_size_bin_kb / 2.0, False)
    out = dict(
        seed=seed, bomb=int(bomb), params8=np.array(list(s.param_simu[0]), dtype=np.float32),
        hist_bins=np.array(s.bins, dtype=np.float64), hist_mean=np.array(s.mean_contacts, dtype=np.float64),
        hist_args=np.array([max_dist_kb, mean_size_bin_kb / 2.0, s.n_frags // 10], dtype=np.float64),
        mean_value_trans_after_fit=np.float64(s.mean_value_trans),
    )
    dumps = dict(A=[], B=[], uniq=[], n_sub=[], structs=[], valid=[])
    orig_eval = s.eval_all_sub_likelihood
    cur = {}

    def spy_eval():
        if len(dumps["A"]) < n_struct_dumps:
            st = np.stack([snapshot(g) for g in s.collector_gpu_vect_frags])
            n_uniq = int(s.gpu_n_uniq.get()[0])
            u = np.full(24, -1, dtype=np.int32)
            u[:n_uniq] = s.gpu_list_uniq_mutations.get()[:n_uniq]
            dumps["A"].append(cur["A"])
            dumps["B"].append(cur["B"])
            dumps["uniq"].append(u)
            dumps["n_sub"].append(int(s.n_sub_vals))
            dumps["structs"].append(st)
            dumps["valid"].append(s.gpu_list_valid_insert.get().copy())
        return orig_eval()

    orig_perform = s.perform_mutations

    def spy_perform(a, b, max_id, is_first):
        cur["A"], cur["B"] = int(a), int(b)
        return orig_perform(a, b, max_id, is_first)

    s.eval_all_sub_likelihood = spy_eval
    s.perform_mutations = spy_perform

    if bomb:
        s.bomb_the_genome()
    out["state0"] = snapshot(s.gpu_vect_frags)
    rec = dict(A=[], cands=[], ncand=[], scores=[], op=[], Bs=[], o=[], dist=[], mean_len=[], n_contigs=[],
               states=[], state_step=[], nuis=[], nuis_step=[], params=[], valid_before=[], params_before=[])
    list_frags = np.arange(0, s.n_new_frags)
    t = 0
    dt = np.float32(0.01)
    for j in range(n_cycles):
        s.gpu_vect_frags.copy_from_gpu()
        np.random.shuffle(list_frags)
        for id_frag in list_frags:
            if t >= max_steps:
                break
            rec["valid_before"].append(s.gpu_list_valid_insert.get().copy())
            rec["params_before"].append(np.array(list(s.param_simu[0]), dtype=np.float32))
            o, dist, op, idb, mean_len, nc = s.step_sampler(id_frag, 5, dt)
            c = np.full(5, -1, dtype=np.int32)
            c[:len(s.candidates)] = s.candidates
            sc = np.zeros(120, dtype=np.float64)
            sc[:len(s.all_scores)] = s.all_scores
            rec["A"].append(int(id_frag)); rec["cands"].append(c); rec["ncand"].append(len(s.candidates))
            rec["scores"].append(sc); rec["op"].append(int(op)); rec["Bs"].append(int(idb)); rec["o"].append(float(o))
            rec["dist"].append(float(dist)); rec["mean_len"].append(np.float32(mean_len)); rec["n_contigs"].append(int(nc))
            if t % state_every == 0:
                rec["states"].append(snapshot(s.gpu_vect_frags)); rec["state_step"].append(t)
            if t >= nuis_after:
                fact, d, d_max, d_nuc, slope, lik, success, _y = s.step_nuisance_parameters(dt, t, n_cycles * s.n_new_frags)
                rec["nuis"].append([float(fact), float(d), float(d_max), float(d_nuc), float(slope),
                                    float(np.ravel(lik)[0]), float(success), float(np.ravel(s.likelihood_nuis)[0])])
                rec["nuis_step"].append(t)
                rec["params"].append(np.array(list(s.param_simu_test[0]), dtype=np.float32))
            t += 1
    out["final_state"] = snapshot(s.gpu_vect_frags)
    for k, v in rec.items():
        out["step_" + k] = np.array(v)
    for k, v in dumps.items():
        out["dump_" + k] = np.array(v)
    return out


CASES = {
    # name: (workload, seed, n_cycles, bomb, max_steps, nuis_after, state_every)
    "micro_seed0": ("micro", 0, 3, False, 100, 70, 1),
    "micro_bomb_seed1": ("micro", 1, 4, True, 150, 110, 1),
    "toy_bomb_seed2": ("toy", 2, 2, True, 260, 200, 1),
}


def main(names=None):
    from instagraal_b200.synth import WORKLOADS, make_level

    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    for name, (wl, seed, ncyc, bomb, max_steps, nuis_after, every) in CASES.items():
        if names and name not in names:
            continue
        level = make_level(WORKLOADS[wl])
        out = run_reference(level, seed, ncyc, bomb, max_steps, nuis_after, state_every=every)
        out["workload"] = wl
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "steps", len(out["step_A"]), "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main(sys.argv[1:] or None)

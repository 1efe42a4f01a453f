#!/usr/bin/env python3
"""Diagnostic: where do the reference's kernels give different results on a real GPU than under the CPU emulation
(oracle/_ref/libref_cpu.so)?  Lockstep product vs reference-cubin on workload T; at a mismatching step the same step is
replayed on the CPU emulation from the same state and the 24 candidate structs of the LAST candidate + all intermediate
vectors are compared field by field."""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="T"); ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--seed", type=int, default=3); ap.add_argument("--max-report", type=int, default=2)
a = ap.parse_args()
from instagraal_b200.synth import make_workload, workload_params
from test_bench_configs import burnt_state
from test_gpu_parity import GpuImpl
from oracle.ref_replay import RefReplaySampler, ALL17
from oracle.sampler_oracle import return_neighbours, setup_distri_frags
level = make_workload(a.workload); p8 = workload_params(level)
st = burnt_state(level, p8, 2, 1000, "bomb")
ref = RefReplaySampler(level, p8, backend="gpu")
ref2 = RefReplaySampler(level, p8, backend="gpu")
cpu = RefReplaySampler(level, p8, backend="cpu")
mine = GpuImpl(level); mine.set_params(p8)
ref.set_state(st)
distri = setup_distri_frags(level.sub_sampled_sparse_matrix, level.n_frags)
frs = np.random.RandomState(a.seed).permutation(level.n_frags)
np.random.seed(a.seed)
state = np.ascontiguousarray(st, dtype=np.int32)
t = rep = 0
for f in frs:
    if t >= a.steps: break
    f = int(f)
    cands = sorted(int(c) for c in return_neighbours(distri, level.n_frags, f, 5) if int(c) != f)
    if not cands: continue
    valid = ref.valid_insert.get().copy()
    mine.set_state(state); mine.set_valid(valid)
    ref.step_sampler(f, cands)
    sa = np.asarray(ref.all_scores, dtype=np.float64)
    sb = np.asarray(mine.step(f, cands)["scores"], dtype=np.float64)
    nz = sa != 0
    rel = float(np.max(np.abs(sa[nz] - sb[nz]) / np.abs(sa[nz])))
    if rel > 1e-9:
        for other, nm in ((ref2, "gpu-again"), (cpu, "cpu-emu")):
            other.set_state(state); other.set_valid(valid)
            other.step_sampler(f, cands)
            so = np.asarray(other.all_scores, dtype=np.float64)
            print("STEP", t, "frag", f, "cands", cands, nm, "max|ref_gpu - %s| =" % nm, float(np.max(np.abs(sa - so))),
                  " max|product - %s| =" % nm, float(np.max(np.abs(sb - so))))
            bad_ops = [g for g in range(len(sa)) if abs(sa[g] - so[g]) > 1e-9 * abs(sa[g])]
            print("   differing proposals (cand*24+op):", bad_ops)
            print("   valid_insert ref:", ref.valid_insert.get().tolist(), nm, other.valid_insert.get().tolist())
            print("   f_up/f_down ref:", ref.f_up.get().tolist(), ref.f_down.get().tolist(), nm, other.f_up.get().tolist(), other.f_down.get().tolist())
            for op in range(24):
                A, B = ref.cand[op].copy_from_gpu(), other.cand[op].copy_from_gpu()
                for k in ALL17:
                    d = np.nonzero(A[k] != B[k])[0]
                    if len(d):
                        print("   last-candidate struct op %2d field %-10s differs at %d frags, e.g. frag %d: gpu %d %s %d" % (op, k, len(d), d[0], A[k][d[0]], nm, B[k][d[0]]))
            for nm2, x, y in (("vect_lik_z", ref.vect_lik_z, other.vect_lik_z), ("all_n_vals_intra", ref.all_n_vals_intra, other.all_n_vals_intra),
                              ("sub_lik_nz", ref.sub_lik_nz, other.sub_lik_nz), ("cur_nz_extract", ref.cur_nz_extract, other.cur_nz_extract),
                              ("cur_nz", ref.cur_nz, other.cur_nz), ("list_uniq", ref.list_uniq, other.list_uniq)):
                print("   ", nm2, "gpu", np.array2string(x.get(), precision=12, max_line_width=250), "\n    ", " " * len(nm2), nm, np.array2string(y.get(), precision=12, max_line_width=250))
        rep += 1
        if rep >= a.max_report: break
    state = ref.get_state()
    t += 1
print("done", t, "steps,", rep, "mismatching steps")

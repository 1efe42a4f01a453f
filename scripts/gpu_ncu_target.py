#!/usr/bin/env python3
"""Lean driver for ncu captures: builds a workload, puts the chain in a state, then runs a few step_sampler calls
(direct launches, no CUDA graph) and a few nuisance likelihood evaluations.
  ncu --set full --clock-control none --import-source on -k regex:k_full_lnz -s 2 -c 1 -f -o gpurun_out/x \\
      python scripts/gpu_ncu_target.py --workload G --state true"""
import argparse, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="G"); ap.add_argument("--state", default="true", choices=["init", "true", "mid", "file"])
ap.add_argument("--state-file", default="/tmp/ig_state13.npy", help="--state file: scaffold saved by an earlier --save-state run (burn-in outside ncu)")
ap.add_argument("--save-state", action="store_true", help="reach the state, save the 13 x NF scaffold to --state-file and exit")
ap.add_argument("--steps", type=int, default=6); ap.add_argument("--nuis", type=int, default=4)
ap.add_argument("--rigid", type=int, default=0); ap.add_argument("--burn", type=int, default=2)
a = ap.parse_args()
from instagraal_b200.synth import make_workload, workload_params
from instagraal_b200.cuda_lib_gl_single import sampler, PARAM_SIMU_RIPPE
level = make_workload(a.workload); p8 = workload_params(level)
s = sampler(*level.sampler_args(), rigid_pruning=a.rigid)
s.set_param_simu(p8)
if a.state == "true":
    s._set_state(level.true_state())
elif a.state == "mid":
    np.random.seed(1000); s.bomb_the_genome(); frs = np.arange(level.n_frags)
    for c in range(a.burn):
        np.random.shuffle(frs); s.run_cycle_device(frs, 5, seed=1000, cycle=c)
elif a.state == "file":
    s._set_state(np.load(a.state_file))
if a.save_state:
    np.save(a.state_file, np.asarray(s._get_state()))
    print("saved", a.state_file, "n_contigs", int(s.n_contigs) if s.n_contigs is not None else None)
    s.free_gpu()
    sys.exit(0)
s.set_options(refresh_every=4096, use_graph=False)
np.random.seed(5)
frs = np.random.permutation(level.n_frags)
for f in frs[:a.steps]:
    s.step_sampler(int(f), 5, np.float32(0.01))
s.param_simu_test = s.param_simu
for i in range(a.nuis):
    p = p8.copy(); p[6] *= np.float32(1.0 + 1e-3 * i)
    s.param_simu_test = np.array([tuple(p.tolist())], dtype=PARAM_SIMU_RIPPE)
    print("nuis", i, s.eval_likelihood_4_nuisance())
print("ok", a.workload, a.state, "n_contigs", int(s.n_contigs) if s.n_contigs is not None else None)
s.free_gpu()

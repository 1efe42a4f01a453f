#!/usr/bin/env python3
"""Times the full-likelihood kernel (k_full_lnz, the nuisance step's evaluate_likelihood_sparse) on a workload.
  python scripts/gpu_r2_full_lnz.py --workload G --state true      (IG_B200_LIB=... selects another build of the library)"""
import argparse, ctypes as C, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="G")
ap.add_argument("--states", default="init,true")
ap.add_argument("--reps", type=int, default=30)
a = ap.parse_args()
from instagraal_b200.synth import make_workload, workload_params
from instagraal_b200.cuda_lib_gl_single import sampler, PARAM_SIMU_RIPPE
from instagraal_b200 import _lib as L
level = make_workload(a.workload)
p8 = workload_params(level)
s = sampler(*level.sampler_args())
s.set_param_simu(p8)
for state in a.states.split(","):
    if state == "true":
        s._set_state(level.true_state())
    s.param_simu_test = s.param_simu
    v0 = s.eval_likelihood()          # fresh coordinates
    vals = []
    has = hasattr(L.lib(), "ig_get_nuisance_stats")
    if has:
        o2 = np.zeros(2); L.lib().ig_get_nuisance_stats(s._h, o2.ctypes.data_as(C.c_void_p), 1)
    t0 = time.perf_counter()
    for i in range(a.reps):
        p = p8.copy(); p[6] *= np.float32(1.0 + 1e-3 * i)
        s.param_simu_test = np.array([tuple(p.tolist())], dtype=PARAM_SIMU_RIPPE)
        vals.append(float(s.eval_likelihood_4_nuisance()))
    wall = (time.perf_counter() - t0) / a.reps * 1e3
    dev = None
    if has:
        L.lib().ig_get_nuisance_stats(s._h, o2.ctypes.data_as(C.c_void_p), 1)
        dev = o2[0] / max(o2[1], 1)
    nnz, ns = s.n_non_zero, int(s.init_n_sub_frags)
    alg = 8 * nnz + 4 * (ns + 1) + 20 * ns
    print(json.dumps(dict(lib=os.environ.get("IG_B200_LIB", "default"), workload=a.workload, state=state, nnz=nnz,
                          lnz_full="%.12e" % v0, nuis_first="%.12e" % vals[0], nuis_last="%.12e" % vals[-1],
                          wall_ms_per_call=wall, device_ms_k_full_lnz=dev, alg_bytes=alg,
                          GBs=(alg / (dev * 1e-3) / 1e9 if dev else None))), flush=True)
s.free_gpu()

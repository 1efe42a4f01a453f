#!/usr/bin/env python3
"""cProfile of the facade's host side: 3000 step_sampler calls, then 1500 step pairs with the nuisance step, workload T."""
import cProfile, io, os, pstats, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from instagraal_b200.synth import make_workload, workload_params
from instagraal_b200.cuda_lib_gl_single import sampler
wl = sys.argv[1] if len(sys.argv) > 1 else "T"
level = make_workload(wl); p8 = workload_params(level)
s = sampler(*level.sampler_args()); s.set_param_simu(p8); s.param_simu_test = s.param_simu
np.random.seed(1); s.bomb_the_genome()
frs = np.arange(level.n_frags)
for c in range(2):
    np.random.shuffle(frs); s.run_cycle_device(frs, 5, seed=1, cycle=c)
dt = np.float32(0.01)
def loop(n, nuis):
    k = 0
    while k < n:
        np.random.shuffle(frs)
        for f in frs:
            s.step_sampler(int(f), 5, dt)
            if nuis: s.step_nuisance_parameters(dt, k, n)
            k += 1
            if k >= n: break
loop(300, False)
for nuis, n in ((False, 3000), (True, 1500)):
    t0 = time.perf_counter(); pr = cProfile.Profile(); pr.enable(); loop(n, nuis); pr.disable(); t = time.perf_counter() - t0
    st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(18)
    print("==== nuisance" if nuis else "==== step only", "ms per iteration (under cProfile):", t / n * 1e3); print(st.getvalue()[:3800])
t0 = time.perf_counter(); loop(3000, False); print("step only, no profiler: ms/step", (time.perf_counter() - t0) / 3000 * 1e3, "device ms/step", s.get_stats(reset=True)["ms_step"] / 3000 if False else "")
t0 = time.perf_counter(); loop(1500, True); print("with nuisance, no profiler: ms/pair", (time.perf_counter() - t0) / 1500 * 1e3)

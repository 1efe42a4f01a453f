#!/usr/bin/env python3
"""bench.py -- headline benchmark of the scaffolding-MCMC hot path (BASELINE.json metric:
"Delta-log-L proposals scored/sec and MCMC cycle wall-time at pyramid level 4").

A *step* is one MCMC step (= one reference ``step_sampler`` call): <=5 candidate neighbours x <=24
moves scored, the best one applied.  Default workload = BASELINE.json configs[1] (yeast-scale
in-silico assembly, level 4, single chain); ``--workload G`` is configs[3] (~1 Gb synthetic).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA through the C ABI)
  python bench.py --impl reference ...                     CPU arm: the oracle port of the reference
                                                           algorithm on all host cores

value   = proposals scored / device time (sum of per-step CUDA-event intervals on the launching
          stream; chain state and contacts resident in HBM; L2 flushed between steps, untimed)
e2e     = same metric through the reference-facing facade (``sampler.step_sampler``): host RNG draw
          of the neighbours, ctypes call, H2D of the candidate list and D2H of the result inside the
          timed region (wall clock).
N > 1   = N independent replica chains (different seeds), one process per GPU, NCCL all-gather of
          {likelihood, n_contigs, live scaffold} every ``--gather-every`` steps; weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

def params_for(level):
    from instagraal_b200.synth import workload_params
    return workload_params(level)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows = []
        self.stop = False
        self.index = index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 7:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(workload, kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel from the
    committed `ncu --set full` capture of the same workload (profiles/r1_ncu_*.txt), or None."""
    table = {("T", "k_score"): 1.300992e6 + 0.255232e6,             # profiles/r1_ncu_k_eval_flat_T.txt + r1_ncu_k_pick_T.txt
             ("G", "k_score"): 171.874048e6 + 9.34016e6,            # profiles/r1_ncu_k_score_G.txt (fully assembled start)
             ("G", "k_full_lnz"): 679.342336e6 + 4.3264e6}          # profiles/r1_ncu_k_full_lnz_G.txt
    return table.get((workload, kernel))


def build_level(name):
    from instagraal_b200.synth import WORKLOADS, make_level, make_workload
    t0 = time.time()
    level = make_workload(name)
    return level, time.time() - t0


def burn_in(s, level, n_cycles, seed):
    """bomb + n_cycles of MCMC so the timed region runs in the assembled (expensive) regime."""
    np.random.seed(seed)
    s.bomb_the_genome()
    frs = np.arange(level.n_frags)
    dt = np.float32(0.01)
    for _ in range(n_cycles):
        np.random.shuffle(frs)
        for f in frs:
            s.step_sampler(int(f), 5, dt)


def run_ours(args):
    import torch  # plumbing only: L2 flush buffer, NCCL all-gather of replica states
    from instagraal_b200.cuda_lib_gl_single import sampler

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)

    level, t_gen = build_level(args.workload)
    p8 = params_for(level)
    s = sampler(*level.sampler_args(), device=local, rigid_pruning=bool(args.rigid_pruning))
    s.set_param_simu(p8)
    burn = args.burn_cycles if args.burn_cycles >= 0 else (2 if level.n_frags <= 5000 else 0)
    t0 = time.time()
    if args.start == "true":  # fully assembled regime without burn-in: one contig per true chromosome
        burn = 0
        np.random.seed(1000 + rank)
        s._set_state(level.true_state())
    elif burn > 0:
        burn_in(s, level, burn, 1000 + rank)
    else:
        np.random.seed(1000 + rank)
        if args.bomb:
            s.bomb_the_genome()
    t_burn = time.time() - t0
    st_burn = s._get_state()
    n_contigs0 = int((st_burn[0] == 0).sum())

    from instagraal_b200.replicas import ReplicaExchange
    xchg = ReplicaExchange(s, dist, dev) if world > 1 else None

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.flush_l2 else None
    dt = np.float32(0.01)
    frs = np.arange(level.n_frags)
    rng = np.random.RandomState(7 + rank)

    def frag_stream(n):
        out = []
        while len(out) < n:
            rng.shuffle(frs)
            out.extend(int(f) for f in frs)
        return out[:n]

    # ---------------- pass 1: device-timed (value, roofline)
    for f in frag_stream(args.warmup):
        s.step_sampler(f, 5, dt)
    s.set_options(refresh_every=args.refresh_every, use_graph=bool(args.graph))
    s.get_stats(reset=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clk:
        t_wall0 = time.perf_counter()
        for i, f in enumerate(frag_stream(args.steps)):
            if flush is not None:
                flush.fill_(i & 0xFF)
                torch.cuda.current_stream().synchronize()   # the flush runs on torch's stream, the step on the library's:
                                                            # it must be over (and stays untimed) before the step starts
            s.step_sampler(f, 5, dt)
            if xchg is not None and (i + 1) % args.gather_every == 0:
                xchg.allgather()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        t_wall = time.perf_counter() - t_wall0
    st = s.get_stats(reset=True)
    dev_ms = st["ms_step"]
    proposals = st["proposals"]

    # ---------------- pass 1b: per-kernel CUDA-event timing for the roofline (profiling on => no graph)
    s.set_profiling(True)
    n_prof = min(args.steps, 500)
    for i, f in enumerate(frag_stream(n_prof)):
        if flush is not None:
            flush.fill_(i & 0xFF)
            torch.cuda.current_stream().synchronize()
        s.step_sampler(f, 5, dt)
    stp = s.get_stats(reset=True)
    ktimes = s.get_kernel_times(reset=True)
    s.set_profiling(False)

    # ---------------- pass 2: end to end through the facade (host RNG + ctypes + H2D/D2H), wall clock
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    np.random.seed(99 + rank)
    t0 = time.perf_counter()
    n_e2e = args.steps
    for f in frag_stream(n_e2e):
        s.step_sampler(f, 5, dt)  # candidates drawn by the facade with np.random (reference contract)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    st2 = s.get_stats(reset=True)

    # ---------------- pass 3: the cycle API (one library call per run of steps, no per-step host sync)
    np.random.seed(199 + rank)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cyc = s.run_cycle(frag_stream(args.steps), 5)
    t_cyc = time.perf_counter() - t0
    cyc_prop = float(cyc["n_proposals"].sum())
    st3 = s.get_stats(reset=True)

    # ---------------- pass 4: the same with the neighbour draws made on the device (production RNG mode)
    s.run_cycle_device(frag_stream(8), 5, seed=7 + rank, cycle=0)   # uploads the neighbour weights, builds graphs
    s.get_stats(reset=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    cycd = s.run_cycle_device(frag_stream(args.steps), 5, seed=7 + rank, cycle=1)
    t_cycd = time.perf_counter() - t0
    cycd_prop = float(cycd["n_proposals"].sum())
    st4 = s.get_stats(reset=True)

    # ---------------- aggregate over ranks (max time, summed proposals)
    vals = np.array([dev_ms, proposals, t_e2e, st2["proposals"], t_wall], dtype=np.float64)
    if dist is not None:
        tv = torch.tensor(vals, device=dev)
        mx = tv.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tv.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms_max, t_e2e_max, t_wall_max = float(mx[0]), float(mx[2]), float(mx[4])
        prop_sum, prop2_sum = float(sm[1]), float(sm[3])
    else:
        dev_ms_max, t_e2e_max, t_wall_max = dev_ms, t_e2e, t_wall
        prop_sum, prop2_sum = proposals, st2["proposals"]

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    value = prop_sum / (dev_ms_max / 1e3)
    e2e = prop2_sum / t_e2e_max
    peak, peak_src = measured_peak()
    nnz, ns = s.n_non_zero, int(s.init_n_sub_frags)
    n_launch_score = stp["steps"]
    n_full = max(stp["full_refreshes"], 0)
    bytes_score = 8 * stp["contacts_read"] + 4 * (stp["rows"] + stp["steps"]) + 16 * stp["rows"] + 32 * stp["frags"]
    bytes_full = (8 * nnz + 4 * (ns + 1) + 20 * ns) * n_full
    kern = {"k_score": (stp["ms_score"], bytes_score, n_launch_score), "k_full_lnz": (stp["ms_full"], bytes_full, n_full)}
    dom = max(kern, key=lambda k: kern[k][0])
    ach = {k: (b / max(ms, 1e-9) / 1e6) for k, (ms, b, _n) in kern.items()}  # GB/s
    out = {
        "metric": "delta-log-L proposals scored per second (MCMC step_sampler, pyramid level 4)",
        "value": value, "unit": "proposals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 expected contacts / f64 log-likelihood accumulation / int32 scaffold",
        "data": "synthetic", "impl": "ours",
        "config": {"workload": args.workload, "n_frags": level.n_frags, "n_sub_frags": ns, "nnz": nnz,
                   "chains": world, "n_neighbours": 5, "start": args.start, "burn_in_cycles": burn, "n_contigs_at_start": n_contigs0,
                   "l2": ("flushed between steps with a 256 MiB write (untimed); timed = sum of per-step CUDA-event "
                          "intervals" if args.flush_l2 else "not flushed (inputs %s L2)" % (">" if nnz * 8 > 126e6 else "<")),
                   "full_likelihood": ("recomputed over every contact each step (reference schedule, CL:1409)"
                                       if args.refresh_every == 1 else
                                       "maintained incrementally, full recompute every %d steps (same values up to f64 "
                                       "summation order; tests/test_gpu_parity.py)" % args.refresh_every),
                   "cuda_graph": bool(args.graph), "rigid_pruning": bool(args.rigid_pruning), "full_refreshes_in_timed_region": st["full_refreshes"],
                   "gather_every": args.gather_every if world > 1 else None},
        "mcmc_cycle_s": dev_ms_max / 1e3 / args.steps * level.n_frags,
        "proposals_per_step": prop_sum / world / args.steps,
        "wall_s_timed_region": t_wall_max,
        "e2e": {"value": e2e, "unit": "proposals/s", "h2d_bytes_per_step": 40, "d2h_bytes_per_step": 6360,  # sizeof(DevScalars): one copy
                "ms_per_step": t_e2e_max / n_e2e * 1e3},
        "e2e_cycle_api": {"value": cyc_prop / t_cyc, "unit": "proposals/s", "ms_per_step": t_cyc / args.steps * 1e3,
                          "device_ms_per_step": st3["ms_step"] / args.steps,
                          "note": "sampler.run_cycle: host draws every step's neighbours (reference RNG order), uploads the "
                                  "plan once, the GPU replays one CUDA graph per step without host synchronisation"},
        "e2e_cycle_device_rng": {"value": cycd_prop / t_cycd, "unit": "proposals/s", "ms_per_step": t_cycd / args.steps * 1e3,
                                 "device_ms_per_step": st4["ms_step"] / args.steps,
                                 "note": "sampler.run_cycle_device: the host uploads the visiting order only; candidates are "
                                         "drawn on the GPU (Philox4x32-10), steps replay as CUDA graphs"},
        "gpu_launches": int(st["launches"]),
        "roofline": {"bound": "hbm", "kernel": dom, "kernel_launches": ("k_pick + k_eval_flat (flat scoring path of small levels)"
                                                                         if dom == "k_score" and level.sparse_matrix.nnz <= 1500000 and level.n_sub_frags <= 16384
                                                                         else dom), "achieved": ach[dom], "peak": peak, "unit": "GB/s",
                     "frac": ach[dom] / peak, "traffic": measured_traffic(args.workload, dom), "peak_source": peak_src,
                     "kernels": {k: {"launches": kern[k][2], "ms_per_launch": kern[k][0] / max(kern[k][2], 1),
                                     "alg_bytes_per_launch": kern[k][1] / max(kern[k][2], 1),
                                     "achieved_GBs": ach[k]} for k in kern},
                     "terms_per_launch": stp["contacts_selected"] * (stp["proposals"] / max(stp["steps"], 1) / 5.0) / max(stp["steps"], 1),
                     "note": "instruction/latency-bound, not HBM-bound: per 8-byte contact up to 24 float32 coordinate comparisons and, where a term changes, powf + f64 log10 (DESIGN.md 4a)"},
        "kernel_us_per_step": {k: v / max(n_launch_score, 1) * 1e3 for k, v in ktimes.items()},
        "clocks": clk.summary(),
        "setup_s": {"generate": t_gen, "burn_in": t_burn},
    }
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline(level, p8, st_burn, args.cpu_budget_s)
    if not args.no_ref_gpu and world == 1:
        try:
            s.free_gpu()
            out["ref_gpu_baseline"] = ref_gpu_baseline(level, p8, st_burn, args.ref_gpu_budget_s, local)
            out["ref_gpu_baseline"]["speedup_e2e_vs_ref_gpu"] = e2e / out["ref_gpu_baseline"]["value"]
        except Exception as ex:  # the baseline must never break the bench line
            out["ref_gpu_baseline"] = {"unavailable": repr(ex)[:200]}
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def _oracle_chain(level, p8, state13, seed, budget_s, max_steps=None, warm=0):
    from oracle.sampler_oracle import OracleSampler
    from instagraal_b200._lib import FIELDS13
    o = OracleSampler(level, p8)
    if state13 is not None:
        o.live = {k: state13[i].copy() for i, k in enumerate(FIELDS13)}
    np.random.seed(seed)
    frs = np.arange(level.n_frags)
    np.random.shuffle(frs)
    for f in frs[:warm]:
        o.step_sampler(int(f), 5)
    t0 = time.perf_counter()
    n_prop = n_steps = 0
    for f in frs[warm:]:
        o.step_sampler(int(f), 5)
        n_prop += int(sum(o.n_uniq_list))
        n_steps += 1
        if time.perf_counter() - t0 > budget_s or (max_steps is not None and n_steps >= max_steps):
            break
    return n_prop, n_steps, time.perf_counter() - t0


def cpu_baseline(level, p8, state13, budget_s):
    """the oracle (NumPy port of the reference algorithm) on one host core, bounded sample"""
    n_prop, n_steps, dt = _oracle_chain(level, p8, state13, 3, budget_s)
    return {"value": n_prop / dt, "unit": "proposals/s", "cores": 1, "kind": "port",
            "sample": "%d step_sampler calls (%d proposals) of the same workload from the same burnt-in scaffold, %.1f s"
                      % (n_steps, n_prop, dt)}


def ref_gpu_baseline(level, p8, state13, budget_s, device=0):
    """B-ref (GPU): the reference's own kernels (oracle/_ref/ref_kernels.cubin) driven in the reference's
    launch order with its per-launch synchronisation, NumPy 'thrust' round trips, 17-array D2H copies
    and Python dist loop (oracle/ref_replay.py, validated against the golden vectors on the CPU backend)."""
    cubin = os.path.join(ROOT, "oracle", "_ref", "ref_kernels.cubin")
    if not os.path.exists(cubin):
        return {"unavailable": "oracle/_ref/ref_kernels.cubin not built"}
    from oracle.ref_replay import RefReplaySampler
    from oracle.sampler_oracle import return_neighbours, setup_distri_frags
    r = RefReplaySampler(level, p8, backend="gpu", device=device)
    if state13 is not None:
        r.set_state(state13)
    distri = setup_distri_frags(level.sub_sampled_sparse_matrix, level.n_frags)
    np.random.seed(3)
    frs = np.arange(level.n_frags)
    np.random.shuffle(frs)
    r.step_sampler(int(frs[0]), return_neighbours(distri, level.n_frags, int(frs[0]), 5))  # warm-up (module load, caches)
    l0 = r.be.n_launch
    t0 = time.perf_counter()
    n_prop = n_steps = 0
    for f in frs[1:]:
        r.step_sampler(int(f), return_neighbours(distri, level.n_frags, int(f), 5))
        n_prop += int(sum(r.n_uniq_list))
        n_steps += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"value": n_prop / dt, "unit": "proposals/s", "ms_per_step": dt / n_steps * 1e3, "steps": n_steps,
            "kernel_launches_per_step": (r.be.n_launch - l0) / n_steps,
            "kind": "reference kernels (cubin built from /root/reference) + restated reference host sequence (route ii)",
            "sample": "%d step_sampler calls from the same burnt-in scaffold, %.1f s" % (n_steps, dt)}


def _ref_worker(a):
    name, seed, budget, max_steps, warm = a
    level, _ = build_level(name)
    p8 = params_for(level)
    return _oracle_chain(level, p8, None, seed, budget, max_steps, warm)


def run_reference(args):
    """CPU arm: the reference's algorithm (oracle port: NumPy transcription of its kernels + its
    orchestration) on every host core, one independent chain per process.  A "step" of this arm is one
    step_sampler call on every core; at most --steps of them are timed after min(--warmup, 2) untimed ones, and
    the sample is cut at --cpu-budget-s seconds so the run stays bounded whatever K is."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    budget = max(5.0, min(args.cpu_budget_s, 60.0))
    level, _ = build_level(args.workload)
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores) as pool:
        warm = max(0, min(args.warmup, 2))
        res = pool.map(_ref_worker, [(args.workload, 100 + i, budget, args.steps, warm) for i in range(cores)])
    wall = time.perf_counter() - t0
    n_prop = sum(r[0] for r in res)
    n_steps = sum(r[1] for r in res)
    t_max = max(r[2] for r in res)
    v = n_prop / t_max
    out = {
        "metric": "delta-log-L proposals scored per second (MCMC step_sampler, pyramid level 4)",
        "value": v, "unit": "proposals/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": max(r[1] for r in res), "warmup": warm, "requested_steps": args.steps, "requested_warmup": args.warmup,
        "ms_per_step": t_max / max(n_steps / cores, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 expected contacts / f64 accumulation / int32 scaffold", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": args.workload, "n_frags": level.n_frags, "n_sub_frags": level.n_sub_frags,
                   "chains": cores, "n_neighbours": 5,
                   "note": "the reference is GPU-only (pycuda); this arm is its algorithm transcribed to NumPy (oracle/), "
                           "one chain per host core from the contig-order start (the in-arm cpu_baseline of the GPU arm "
                           "times the same port on the GPU arm's burnt-in scaffold: same rate per core)"},
        "cpu_baseline": {"value": v, "unit": "proposals/s", "cores": cores, "kind": "port",
                         "sample": "%d processes x %.0f s of step_sampler calls (%d steps, %d proposals), wall %.1f s"
                                   % (cores, budget, n_steps, n_prop, wall)},
        "e2e": {"value": v, "unit": "proposals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8000)
    ap.add_argument("--warmup", type=int, default=500)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="T")
    ap.add_argument("--burn-cycles", type=int, default=-1)
    ap.add_argument("--bomb", type=int, default=1)
    ap.add_argument("--start", default="bomb", choices=["bomb", "true"])
    ap.add_argument("--flush-l2", type=int, default=1)
    ap.add_argument("--gather-every", type=int, default=500)
    ap.add_argument("--refresh-every", type=int, default=4096)
    ap.add_argument("--graph", type=int, default=1)
    ap.add_argument("--rigid-pruning", type=int, default=0,
                    help="1: skip contacts whose ends move rigidly together (not the reference's float32 re-rounding; see DESIGN.md)")
    ap.add_argument("--cpu-budget-s", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--ref-gpu-budget-s", type=float, default=10.0)
    args = ap.parse_args()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    if args.impl == "reference":
        run_reference(args)
    else:
        import __graft_entry__ as g
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            g.build()
        run_ours(args)


if __name__ == "__main__":
    main()

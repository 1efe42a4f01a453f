"""On-device timeline of the step's kernels inside CUDA-graph replay with warm caches (cycle API).
Needs the -DIG_TIMELINE build:  IG_B200_LIB=instagraal_b200/libinstagraal_b200_tl.so python scripts/timeline.py [workload] [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from instagraal_b200 import _lib as L  # noqa: E402
from instagraal_b200.cuda_lib_gl_single import sampler  # noqa: E402
from instagraal_b200.synth import make_workload  # noqa: E402

NAMES = ["cand_setup", "find_cuts", "classes", "rows", "rows_write", "precompute", "score", "finalize", "lnz_outside", "apply",
         "post", "commit_coords", "prefetch", "coords", "full_lnz", "pick"]
MAIN = ["commit_coords", "cand_setup", "find_cuts", "rows", "rows_write", "precompute", "pick", "score", "finalize", "apply", "post"]


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "T"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    level = make_workload(wl)
    s = sampler(*level.sampler_args())
    p8 = np.array([50.0, 9.6, np.float32(0.53 * (9.6 / 50.0) ** -1.5 * 50.0 ** -3), -1.5, 2.0, 900.0, 4.0e5, 0.02], dtype=np.float32)
    s.set_param_simu(p8)
    np.random.seed(1)
    s.bomb_the_genome()
    rng = np.random.RandomState(0)
    for c in range(2):   # burn-in like bench.py
        s.run_cycle_device(rng.permutation(level.n_frags), 5, seed=1, cycle=c)
    frs = np.concatenate([rng.permutation(level.n_frags) for _ in range(1 + n // level.n_frags)])[:n]
    L.check(s._h, L.lib().ig_timeline_reset(s._h), "ig_timeline_reset")
    ph = np.zeros(16, dtype=np.uint64)
    L.check(s._h, L.lib().ig_timeline_phases(s._h, ph.ctypes.data, 1), "ig_timeline_phases")
    s.run_cycle_device(frs, 5, seed=1, cycle=7)
    L.check(s._h, L.lib().ig_timeline_phases(s._h, ph.ctypes.data, 0), "ig_timeline_phases")
    tl = np.zeros((n, 16, 2), dtype=np.uint64)
    L.check(s._h, L.lib().ig_timeline_get(s._h, n, tl.ctypes.data), "ig_timeline_get")
    tl = tl[50:].astype(np.float64)            # skip the first steps (full refresh, cold)
    ran = tl[:, :, 1] > 0
    dur = np.where(ran, tl[:, :, 1] - tl[:, :, 0], np.nan) / 1e3
    t0 = np.nanmin(np.where(ran, tl[:, :, 0], np.nan), axis=1)
    t1 = np.nanmax(np.where(ran, tl[:, :, 1], np.nan), axis=1)
    period = np.diff(t0) / 1e3
    print("workload %s, %d steps of run_cycle_device (CUDA-graph replay, back to back, warm L2)" % (wl, len(tl)))
    print("step period: mean %.1f us, median %.1f us; first-start to last-end inside a step: mean %.1f us" %
          (period.mean(), np.median(period), ((t1 - t0) / 1e3).mean()))
    print("%-14s %8s %8s   %s" % ("kernel", "mean us", "median", "start offset in step (mean us)"))
    for i, nm in enumerate(NAMES):
        if ran[:, i].any():
            off = np.nanmean(np.where(ran[:, i], tl[:, i, 0] - t0, np.nan)) / 1e3
            print("%-14s %8.2f %8.2f   %8.2f" % (nm, np.nanmean(dur[:, i]), np.nanmedian(dur[:, i]), off))
    prev = None
    gaps = []
    for nm in MAIN:
        i = NAMES.index(nm)
        if not ran[:, i].any():
            continue
        if prev is not None:
            g = (tl[:, i, 0] - tl[:, prev, 1]) / 1e3
            gaps.append((NAMES[prev] + " -> " + nm, np.nanmean(np.where(ran[:, i] & ran[:, prev], g, np.nan))))
        prev = i
    print("gaps on the main chain (end of one kernel to first block of the next):")
    for k, v in gaps:
        print("  %-28s %6.2f us" % (k, v))
    print("  sum of gaps %.1f us; inter-step gap (post end -> next step first start): %.2f us" %
          (sum(v for _, v in gaps), np.mean((t0[1:] - t1[:-1]) / 1e3)))
    tot = max(float(ph[:5].sum()), 1.0)
    print("k_score phases (cycles of warp 0 of every block, share): prologue %.1f%%, item set-up %.1f%%, contact loop %.1f%%, final flush + reductions %.1f%%, block barrier wait %.1f%%; mean cycles per block %.0f" %
          tuple([100 * float(x) / tot for x in ph[:5]] + [tot / max(n, 1) / 2220.0]))
    if ph[8:15].sum() > 0:
        tp, te = float(ph[8:12].sum()), float(ph[12:15].sum())
        print("k_pick phases (thread 0 of every block, share of %.0f cycles per step): prologue %.1f%%, chunk loop %.1f%%, barrier+ticket %.1f%%, offsets scan (last block) %.1f%%" %
              ((tp / n,) + tuple(100 * float(x) / tp for x in ph[8:12])))
        print("k_eval_flat phases (share of %.0f cycles per step): prologue %.1f%%, items %.1f%%, final flush + reductions %.1f%%" %
              ((te / n,) + tuple(100 * float(x) / te for x in ph[12:15])))
    tf = float(ph[5:8].sum())
    print("k_finalize phases (thread 0 of every block): partial reductions %.1f%%, last-block quirk %.1f%%, scores %.1f%%; mean cycles per block %.0f" %
          (100 * float(ph[5]) / tf, 100 * float(ph[6]) / tf, 100 * float(ph[7]) / tf, tf / max(n, 1) / 5.0))
    # per-block trace of the last k_score launch
    nb = 8192
    tb = np.zeros((nb, 4), dtype=np.uint64)
    L.check(s._h, L.lib().ig_timeline_blocks(s._h, nb, tb.ctypes.data), "ig_timeline_blocks")
    tb = tb[tb[:, 1] > 0].astype(np.float64)
    if len(tb):
        t00 = tb[:, 0].min()
        st, en = (tb[:, 0] - t00) / 1e3, (tb[:, 1] - t00) / 1e3
        d = en - st
        busy = tb[:, 3] > 0
        print("last k_score launch: %d blocks (%d with work), kernel span %.1f us" % (len(tb), busy.sum(), en.max()))
        print("  block duration us: all mean %.2f p50 %.2f p90 %.2f max %.2f | with work mean %.2f max %.2f | idle blocks mean %.2f" %
              (d.mean(), np.median(d), np.percentile(d, 90), d.max(), d[busy].mean(), d[busy].max(), d[~busy].mean() if (~busy).any() else 0))
        print("  block start us: p10 %.1f p50 %.1f p90 %.1f max %.1f ; items per busy block mean %.1f max %d" %
              (np.percentile(st, 10), np.median(st), np.percentile(st, 90), st.max(), tb[busy, 3].mean(), tb[:, 3].max()))
        sm = tb[:, 2].astype(int)
        per_sm = np.bincount(sm)
        print("  blocks per SM: min %d max %d ; SM-busy time (sum of block durations / 3 slots) mean %.1f us max %.1f us" %
              (per_sm[per_sm > 0].min(), per_sm.max(), np.bincount(sm, weights=d).mean() / 3, np.bincount(sm, weights=d).max() / 3))
        # concurrency over time
        ts = np.linspace(0, en.max(), 40)
        conc = [(int(((st <= t) & (en > t)).sum())) for t in ts]
        print("  resident blocks over time:", conc)
    s.free_gpu()


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE ONLY -- restatement of the product's production-mode neighbour draws
(instagraal_b200/csrc/ig_k_rng.cuh: philox4x32_10, philox_uniform, k_draw_plan) in NumPy/Python.

The distribution is the reference's return_neighbours (cuda_lib_gl_single.py:3103-3141: min(delta, #non-zero pk)
fragments without replacement, probability proportional to pk; `delta` distinct uniform fragments when the
fragment has no neighbour), then sorted (cuda_lib_gl_single.py:1404); the random stream is Philox4x32-10
(Salmon et al. 2011, the counter-based generator of Random123 / cuRAND) instead of NumPy's MT19937.
Only tests import this module."""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(c, k):
    c = list(c)
    k = list(k)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k[1]) & MASK, p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return c


def uniform(step, cycle, draw, attempt, seed):
    r = philox4x32_10((step, cycle, draw, attempt), (seed & MASK, (seed >> 32) & MASK))
    bits = (r[0] << 32) | r[1]
    return float(bits >> 11) * (1.0 / 9007199254740992.0)


def draw_plan(frags, delta, n_frags, ptr, idx, cdf, n_nonzero, seed, cycle, max_cands=8):
    plan = np.zeros((len(frags), 2 + max_cands), dtype=np.int32)
    for t, a in enumerate(frags):
        a = int(a)
        b, e = int(ptr[a]), int(ptr[a + 1])
        got = []
        if e > b:
            n_max = min(delta, int(n_nonzero[a]))
            total = float(cdf[e - 1])
            row = cdf[b:e]
            for i in range(n_max):
                pick = -1
                for att in range(256):
                    u = uniform(t, cycle, i, att, seed) * total
                    j = int(np.searchsorted(row, u, side="right"))   # first j with cdf[j] > u
                    j = min(j, e - b - 1)
                    c = int(idx[b + j])
                    if c not in got:
                        pick = c
                        break
                if pick < 0:
                    w = np.diff(np.concatenate([[0.0], row]))
                    for j in range(e - b):
                        if w[j] > 0 and int(idx[b + j]) not in got:
                            pick = int(idx[b + j])
                            break
                if pick >= 0:
                    got.append(pick)
        else:
            for i in range(min(delta, n_frags - 1)):
                for att in range(256):
                    c = min(n_frags - 1, int(uniform(t, cycle, i, att, seed) * float(n_frags)))
                    if c != a and c not in got:
                        got.append(c)
                        break
        got = sorted(g for g in got if g != a)
        plan[t, 0] = len(got)
        plan[t, 1] = a
        plan[t, 2:2 + len(got)] = got
    return plan

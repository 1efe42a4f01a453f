"""Host RNG contract helpers.

``weighted_choice_no_replace`` returns exactly what ``np.random.choice(a, size, p=p, replace=False)`` of the legacy
global ``RandomState`` returns and leaves the generator in exactly the same state, but without that call's per-call
validation of ``p`` (finite, non-negative, sums to 1 by Kahan summation, ...) and re-computation of its cumulative
sum -- which dominate the cost of ``return_neighbours`` (cuda_lib_gl_single.py:3103-3141 in the reference) once
the GPU part of a step takes ~0.1 ms.  NumPy's algorithm (numpy/random/mtrand.pyx, ``choice``, branch
``replace=False`` with ``p``): draw ``size - n_found`` uniforms, look them up in the normalised cumulative sum of ``p``
with the already found entries zeroed, keep the first occurrences in draw order, repeat until ``size`` are found.
The first round uses a cumulative sum precomputed once per fragment; only when it yields duplicates is NumPy's loop
continued literally.  Checked against ``np.random.choice`` draw for draw, state for state (tests/test_host_rng.py).
"""
import numpy as np


def prepare(pk):
    """per-fragment constants: p as float64 (what mtrand converts it to) and its normalised cumulative sum"""
    p64 = np.ascontiguousarray(pk, dtype=np.float64)
    cdf = np.cumsum(p64)
    cdf /= cdf[-1]
    return p64, cdf


def weighted_choice_no_replace(a, p64, cdf0, size):
    if size <= 0:
        return a[:0]
    x = np.random.random_sample(size)
    new = cdf0.searchsorted(x, side="right")
    if size == 1 or len(set(new.tolist())) == size:
        return a[new]
    p = p64.copy()
    found = np.zeros(size, dtype=np.int64)
    _, ui = np.unique(new, return_index=True)
    ui.sort()
    new = new.take(ui)
    n_uniq = new.size
    found[:n_uniq] = new
    while n_uniq < size:
        x = np.random.random_sample(size - n_uniq)
        p[found[0:n_uniq]] = 0
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        new = cdf.searchsorted(x, side="right")
        _, ui = np.unique(new, return_index=True)
        ui.sort()
        new = new.take(ui)
        found[n_uniq:n_uniq + new.size] = new
        n_uniq += new.size
    return a[found]

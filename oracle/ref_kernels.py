"""TEST INFRASTRUCTURE ONLY: direct ctypes access to the reference's own kernels compiled for the
CPU (oracle/_ref/libref_cpu.so, built in place from /root/reference by oracle/Makefile).

Used by the fuzz tests to compare the oracle restatement (oracle/moves.py) -- and through it the
CUDA product -- against the REAL reference kernels on random scaffolds, including circular
contigs that the recorded trajectories never reach.  The launch sequence restated here is
perform_mutations (cuda_lib_gl_single.py:1642-1923); the kernels themselves are the reference's.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .moves import ALL17, FIELDS

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_cpu.so")


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.emu_launch.restype = ctypes.c_int
        _lib.emu_launch.argtypes = [
            ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint,
            ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
    return _lib


class Struct:
    """A ``frag`` struct (KA:40-58): 17 int32 arrays + the packed array of 17 pointers."""

    def __init__(self, n, state=None, fill=0):
        self.n = n
        self.a = {k: np.full(n, fill, dtype=np.int32) for k in ALL17}
        if state is not None:
            self.load(state)
        self.ptrs = np.array([self.a[k].ctypes.data for k in ALL17], dtype=np.uint64)

    def load(self, state):
        n = self.n
        for k in FIELDS:
            self.a[k][:] = state[k]
        self.a["id"][:] = np.arange(n)
        self.a["rep"][:] = 0
        self.a["activ"][:] = 1
        self.a["id_d"][:] = np.arange(n)

    def ptr(self):
        return self.ptrs.ctypes.data

    def state(self):
        return {k: self.a[k].copy() for k in FIELDS}


def launch(name, n_threads_total, block, *int_args):
    L = lib()
    fn = ctypes.cast(getattr(L, name), ctypes.c_void_p).value
    grid = n_threads_total // block + 1
    ia = (ctypes.c_uint64 * len(int_args))(*[int(x) & 0xFFFFFFFFFFFFFFFF for x in int_args])
    fa = (ctypes.c_uint32 * 1)(0)
    rc = L.emu_launch(fn, grid, 1, block, 1, 1, len(int_args), ia, 0, fa)
    assert rc == 0, (name, rc)


def perform_mutations(state, a, b, max_id, sentinel=-777):
    """Reference kernels, reference launch order.  Returns (list of 24 states, valid[12],
    stale_hit: True when some output entry was never written = quirk Q4 reached)."""
    n = len(state["pos"])
    live = Struct(n, state)
    out = [Struct(n, fill=sentinel) for _ in range(24)]
    pop = Struct(n)
    t1 = Struct(n)
    t2 = Struct(n)
    ids = np.zeros(n, dtype=np.int32)
    idp = ids.ctypes.data
    for mode in range(8):
        launch("pop_out_frag", n, 1024, pop.ptr(), live.ptr(), idp, a, max_id, n)
        max_id2 = int(ids.max())
        if mode == 0:
            launch("simple_copy", n, 1024, out[0].ptr(), pop.ptr(), n)
        elif mode == 1:
            launch("flip_frag", n, 1024, out[1].ptr(), live.ptr(), a, n)
        else:
            kern = ("pop_in_frag_1", "pop_in_frag_1", "pop_in_frag_2", "pop_in_frag_2", "pop_in_frag_3",
                    "pop_in_frag_3")[mode - 2]
            launch(kern, n, 1024, out[mode].ptr(), pop.ptr(), a, b, max_id2, 1 if mode % 2 == 0 else -1, n)
    mode = 0
    for up_a in (0, 1):
        launch("split_contig", n, 128, t1.ptr(), live.ptr(), idp, a, up_a, max_id, n)
        for up_b in (0, 1):
            launch("split_contig", n, 128, t1.ptr(), live.ptr(), idp, a, up_a, max_id, n)
            max_id1 = int(ids.max())
            launch("split_contig", n, 128, t2.ptr(), t1.ptr(), idp, b, up_b, max_id1, n)
            max_id2 = int(ids.max())
            launch("paste_contigs", n, 128, out[8 + mode].ptr(), t2.ptr(), a, b, max_id2, n)
            mode += 1
    valid = np.full(12, -1, dtype=np.int32)
    bounds = np.array([1, 3, 5, 10, 20, 50], dtype=np.int32)
    f_up = np.full(6, -1, dtype=np.int32)
    f_down = np.full(6, -1, dtype=np.int32)
    launch("get_bounds", n, 64, live.ptr(), a, b, valid.ctypes.data, bounds.ctypes.data, f_up.ctypes.data,
           f_down.ctypes.data, 6, n)
    k = 0
    for i in range(6):
        for j in (1, 0):
            lst = f_up if j == 1 else f_down
            launch("extract_block", n, 64, t1.ptr(), live.ptr(), idp, a, lst.ctypes.data, i, j, max_id, n)
            launch("insert_block", n, 64, out[12 + k].ptr(), t1.ptr(), live.ptr(), a, b, lst.ctypes.data,
                   valid.ctypes.data, k, i, j, n)
            k += 1
    stale = any((o.a[f] == sentinel).any() for o in out for f in FIELDS)
    return [o.state() for o in out], valid.tolist(), stale

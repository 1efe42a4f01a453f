"""TEST / BASELINE INFRASTRUCTURE ONLY (never imported by the product path).

Replays the reference's ``step_sampler`` launch sequence (cuda_lib_gl_single.py:1401-1465 and
everything it calls, "CL") on the reference's OWN kernels:

  * backend "cpu": oracle/_ref/libref_cpu.so  (kernel_sparse_adapt.cu compiled in place for the CPU)
  * backend "gpu": oracle/_ref/ref_kernels.cubin (same source, nvcc -arch=sm_100a) through the CUDA
    driver API of cuda-python -- this is baseline "B-ref (GPU), route (ii)" of BASELINE.md: the
    reference kernels, the reference launch order, a blocking synchronise after every launch like
    the reference's ``end.synchronize()``, the NumPy "thrust" round trips (CL:28-88), the 17-array
    D2H copies (CL:1410, 2095, 666) and the Python loop of dist_inter_genome (CL:665-716).

pycuda itself is not installable in this image and the reference's Python cannot travel to the GPU
box, hence this restatement of the *host* sequence; it is validated on the CPU backend against the
golden vectors recorded from the unmodified reference class (tests/test_ref_replay.py).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from .moves import ALL17, FIELDS
from .sampler_oracle import dist_inter_genome, upper_coo

_HERE = os.path.dirname(os.path.abspath(__file__))
N_TMP = 24


# ------------------------------------------------------------------------------------------ backends
class CpuBackend:
    name = "cpu"

    def __init__(self):
        self.lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "libref_cpu.so"))
        self.lib.emu_launch.restype = ctypes.c_int
        self.lib.emu_launch.argtypes = [ctypes.c_void_p] + [ctypes.c_uint] * 5 + [
            ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.POINTER(ctypes.c_uint32)]
        self._keep = []
        self.n_launch = 0

    def alloc(self, nbytes):
        buf = np.zeros(max(int(nbytes), 8) + 64, dtype=np.uint8)
        self._keep.append(buf)
        return buf.ctypes.data

    def h2d(self, ptr, arr):
        a = np.ascontiguousarray(arr)
        ctypes.memmove(ptr, a.ctypes.data, a.nbytes)

    def d2h(self, arr, ptr):
        ctypes.memmove(arr.ctypes.data, ptr, arr.nbytes)

    def launch(self, name, grid, block, args):
        ints = [int(v) & 0xFFFFFFFFFFFFFFFF for k, v in args if k not in ("f", "x")]
        fps = [int(np.float32(v).view(np.uint32)) for k, v in args if k == "f"]
        ia = (ctypes.c_uint64 * max(len(ints), 1))(*ints)
        fa = (ctypes.c_uint32 * max(len(fps), 1))(*fps)
        fn = ctypes.cast(getattr(self.lib, name), ctypes.c_void_p).value
        rc = self.lib.emu_launch(fn, int(grid), 1, int(block), 1, 1, len(ints), ia, len(fps), fa)
        assert rc == 0, (name, rc)
        self.n_launch += 1

    def sync(self):
        pass


class GpuBackend:
    name = "gpu"

    def __init__(self, device=0):
        from cuda.bindings import driver as drv
        self.drv = drv
        self._ck(drv.cuInit(0))
        dev = self._ck(drv.cuDeviceGet(device))
        self.ctx = self._ck(drv.cuDevicePrimaryCtxRetain(dev))
        self._ck(drv.cuCtxSetCurrent(self.ctx))
        data = open(os.path.join(_HERE, "_ref", "ref_kernels.cubin"), "rb").read()
        self.mod = self._ck(drv.cuModuleLoadData(data))
        self.fn = {}
        self.n_launch = 0

    def _ck(self, res):
        err = res[0]
        if int(err) != 0:
            raise RuntimeError("CUDA driver error %s" % err)
        return res[1] if len(res) == 2 else res[1:]

    def alloc(self, nbytes):
        p = self._ck(self.drv.cuMemAlloc(max(int(nbytes), 8)))
        self._ck(self.drv.cuMemsetD8(p, 0, max(int(nbytes), 8)))
        return int(p)

    def h2d(self, ptr, arr):
        a = np.ascontiguousarray(arr)
        self._ck(self.drv.cuMemcpyHtoD(ptr, a.ctypes.data, a.nbytes))

    def d2h(self, arr, ptr):
        self._ck(self.drv.cuMemcpyDtoH(arr.ctypes.data, ptr, arr.nbytes))

    def launch(self, name, grid, block, args):
        f = self.fn.get(name)
        if f is None:
            f = self.fn[name] = self._ck(self.drv.cuModuleGetFunction(self.mod, name.encode()))
        vals, types = [], []
        for k, v in args:
            if k == "p":
                vals.append(int(v)); types.append(ctypes.c_void_p)
            elif k == "f":
                vals.append(float(v)); types.append(ctypes.c_float)
            elif k == "q":
                vals.append(int(v)); types.append(ctypes.c_ulonglong)
            elif k == "x":  # unused float2 parameter of flip_frag (KA:612-613): the driver checks the count
                vals.append(0.0); types.append(ctypes.c_double)
            else:
                vals.append(int(v)); types.append(ctypes.c_int)
        self._ck(self.drv.cuLaunchKernel(f, int(grid), 1, 1, int(block), 1, 1, 0, 0, (tuple(vals), tuple(types)), 0))
        self._ck(self.drv.cuCtxSynchronize())  # the reference records an event and synchronises after every launch
        self.n_launch += 1

    def sync(self):
        self._ck(self.drv.cuCtxSynchronize())


class DevArray:
    def __init__(self, be, arr):
        arr = np.ascontiguousarray(arr)
        self.be, self.dtype, self.shape, self.nbytes = be, arr.dtype, arr.shape, arr.nbytes
        self.ptr = be.alloc(arr.nbytes)
        be.h2d(self.ptr, arr)

    def get(self):
        out = np.empty(self.shape, dtype=self.dtype)
        self.be.d2h(out, self.ptr)
        return out

    def set(self, arr):
        self.be.h2d(self.ptr, np.ascontiguousarray(arr, dtype=self.dtype))

    def fill(self, v):
        self.set(np.full(self.shape, v, dtype=self.dtype))


class DevStruct:
    """GPUStruct (gpustruct.py): 17 int32 arrays + the device-resident array of 17 pointers."""

    def __init__(self, be, n, state=None):
        self.be, self.n = be, n
        self.arr = {}
        for k in ALL17:
            a = np.zeros(n, dtype=np.int32)
            if state is not None and k in state:
                a[:] = state[k]
            elif state is not None and k in ("id", "id_d"):
                a[:] = np.arange(n)
            elif state is not None and k == "activ":
                a[:] = 1
            self.arr[k] = DevArray(be, a)
        self.ptrs = DevArray(be, np.array([self.arr[k].ptr for k in ALL17], dtype=np.uint64))
        self.ptr = self.ptrs.ptr

    def copy_from_gpu(self):
        return {k: self.arr[k].get() for k in ALL17}


# ------------------------------------------------------------------------------------------ the replay
class RefReplaySampler:
    def __init__(self, level, params8, backend="cpu", device=0, canonical_slice_order=False):
        """canonical_slice_order: slice_sp_mat (KA:485-607) appends the selected contacts through an atomic counter, so
        their order inside a row is whatever order the atomics executed in, and the stable host sort by row (CL:45-55)
        keeps it.  Under the CPU emulation that order is ascending column; on a real GPU it VARIES FROM RUN TO RUN, and
        eval_sub_likelihood's last-block quirk (KA:4362: uniq positions >= n_sub % 64 lose the final block's contacts)
        makes the scores of those positions depend on it (two replays of one step from one state on the same B200
        differ, profiles/r2_reference_nondeterminism.txt).  True sorts by (row, column) instead: the reference's kernels
        stay untouched, its one arbitrary choice is pinned to the emulation's."""
        self.canonical_slice_order = canonical_slice_order
        be = self.be = CpuBackend() if backend == "cpu" else GpuBackend(device)
        self.nf, self.ns = level.n_frags, level.n_sub_frags
        nf, ns = self.nf, self.ns
        rows, cols, dat = upper_coo(level.sparse_matrix)
        self.nnz = len(dat)
        self.sp_dat, self.sp_rows, self.sp_cols = DevArray(be, dat), DevArray(be, rows), DevArray(be, cols)
        self.sub_dat = DevArray(be, np.zeros_like(dat))
        self.sub_rows = DevArray(be, np.zeros_like(dat))
        self.sub_cols = DevArray(be, np.zeros_like(dat))
        self.s2f = DevArray(be, level.np_sub_frags_2_frags)
        init = {k: np.asarray(level.S_o_A_frags[k], dtype=np.int32) for k in FIELDS if k != "ori"}
        init["ori"] = np.ones(nf, dtype=np.int32)
        self.live = DevStruct(be, nf, init)
        self.cand = [DevStruct(be, nf) for _ in range(N_TMP)]
        self.pop, self.t1, self.t2 = DevStruct(be, nf), DevStruct(be, nf), DevStruct(be, nf)
        z = np.zeros(nf, dtype=np.int32)
        self.id_contigs, self.pop_ids, self.t1_ids, self.t2_ids = (DevArray(be, z) for _ in range(4))
        self.collect = [DevArray(be, np.ones(ns * N_TMP, dtype=dt)) for dt in (np.float32, np.int32, np.float32, np.int32, np.int32)]
        self.vect = [DevArray(be, np.ones(ns, dtype=dt)) for dt in (np.float32, np.int32, np.float32, np.int32, np.int32)]
        ident = np.arange(nf, dtype=np.int32)
        self.collector_id = DevArray(be, ident)
        self.dispatcher = DevArray(be, np.stack([ident, ident + 1], axis=1).astype(np.int32))
        sident = np.arange(ns, dtype=np.int32)
        self.sub_collector_id = DevArray(be, sident)
        self.sub_dispatcher = DevArray(be, np.stack([sident, sident + 1], axis=1).astype(np.int32))
        self.id_single = DevArray(be, sident)
        self.params = DevArray(be, np.asarray(params8, dtype=np.float32))
        self.mbar = np.float32(level.S_o_A_sub_frags["len_bp"].mean() / 1000.0)
        self.mbar_q1 = np.int32(level.S_o_A_sub_frags["len_bp"].mean() / 1000.0)  # CL:743 passes an int (quirk Q1)
        with np.errstate(over="ignore"):
            ns32 = np.int32(ns)
            self.n_pix = np.float64(ns32 * (ns32 - np.int32(1)) / 2)
        self.max_bounds_insert = int(50 * np.int32(np.round(level.S_o_A_frags["sub_len"].mean()) + 1))
        d1 = lambda dt, n=1: DevArray(be, np.zeros(n, dtype=dt))
        self.counter, self.counter_gl = d1(np.int32), d1(np.int32)
        self.lik_zeros, self.vect_lik_z = d1(np.float64), d1(np.float64, N_TMP)
        self.n_vals_intra, self.all_n_vals_intra = d1(np.int32), d1(np.int32, N_TMP)
        self.list_uniq, self.n_uniq = d1(np.int32, N_TMP), d1(np.int32)
        self.n_pix_dev = DevArray(be, np.array([self.n_pix], dtype=np.float64))
        self.block_indptr = d1(np.int32, max(self.nnz, 1))
        self.info_blocks = d1(np.int32, 3 * (self.nnz // 64 + 2))
        self.sub_lik_nz, self.cur_nz_extract = d1(np.float64, N_TMP), d1(np.float64)
        self.all_scores_dev, self.cur_nz = d1(np.float64, N_TMP), d1(np.float64)
        self.uniq_id_c, self.uniq_len = d1(np.int32, ns), d1(np.int32, ns)
        self.old2new = d1(np.int32, ns + ns // 10 + 2)
        self.valid_insert = d1(np.int32, 12)
        self.list_bounds = DevArray(be, np.array([1, 3, 5, 10, 20, 50], dtype=np.int32))
        self.f_up, self.f_down = d1(np.int32, 6), d1(np.int32, 6)
        self.gl_pos, self.gl_vel = d1(np.float32, 4 * nf), d1(np.float32, 4 * nf)
        self.gl_pos_gen, self.gl_vel_gen = d1(np.float32, 4 * nf), d1(np.float32, 4 * nf)
        self.rng = d1(np.uint8, 100 * 48)
        be.launch("init_rng", 100 // 64 + 1, 64, [("i", 100), ("p", self.rng.ptr), ("q", 1), ("q", 0)])
        self.init_prev = np.copy(level.S_o_A_frags["prev"])
        self.init_next = np.copy(level.S_o_A_frags["next"])
        self.orientable = (level.np_sub_frags_id["w"] > 1).astype(np.int32)
        self.host = None

    # --- helpers mirroring the reference methods
    def _L(self, name, n, block, args):
        self.be.launch(name, int(n) // block + 1, block, args)

    def copy_from_gpu(self):
        self.host = self.live.copy_from_gpu()
        return self.host

    def fill_dist_single(self):  # CL:936-970
        a = [("p", self.s2f.ptr), ("p", self.live.ptr)] + [("p", v.ptr) for v in self.vect] + [
            ("p", self.collector_id.ptr), ("p", self.dispatcher.ptr), ("p", self.sub_collector_id.ptr),
            ("p", self.sub_dispatcher.ptr), ("i", self.ns)]
        self._L("uni_fill_vect_dist", self.ns, 1024, a)

    def eval_likelihood(self):  # CL:1245-1294
        self.fill_dist_single()
        self.lik_zeros.fill(0); self.n_vals_intra.fill(0)
        v = self.vect
        self._L("eval_likelihood_on_zero", self.ns, 1024, [("p", v[1].ptr), ("p", v[2].ptr), ("p", v[3].ptr), ("p", v[4].ptr),
                ("p", self.params.ptr), ("f", np.int32(self.mbar_q1).view(np.float32)), ("p", self.lik_zeros.ptr),
                ("p", self.n_vals_intra.ptr), ("i", self.ns)])
        self.lik_zeros.get(); self.n_vals_intra.get()
        self.cur_nz.fill(0.0)
        self._L("evaluate_likelihood_sparse", self.nnz, 1024, [("p", self.sp_dat.ptr), ("p", self.sp_rows.ptr), ("p", self.sp_cols.ptr),
                ("p", self.id_single.ptr), ("p", self.params.ptr), ("f", self.mbar)] + [("p", x.ptr) for x in v] + [
                ("p", self.cur_nz.ptr), ("i", self.nnz), ("i", self.ns)])

    def modify_gl_cuda_buffer(self, id_fi):  # CL:2715-2881
        nf = self.nf
        self.counter_gl.fill(0)
        self.be.launch("select_uniq_id_c", int(nf / 512 + 1), 512, [("p", self.live.ptr), ("p", self.uniq_id_c.ptr),
                       ("p", self.uniq_len.ptr), ("p", self.counter_gl.ptr), ("i", nf)])
        nc = int(self.counter_gl.get()[0])
        keys, vals = self.uniq_len.get(), self.uniq_id_c.get()  # the NumPy "thrust" round trip (CL:69-77)
        idx = np.argsort(-keys[:nc].astype(np.int64), kind="stable")
        keys[:nc], vals[:nc] = keys[:nc][idx], vals[:nc][idx]
        self.uniq_len.set(keys); self.uniq_id_c.set(vals)
        lens = np.float32(self.uniq_len.get())
        self.n_contigs = np.int32(nc)
        self.mean_length_contigs = lens[:nc].mean()
        self.be.launch("make_old_2_new_id_c", int(nc / 512 + 1), 512, [("p", self.uniq_id_c.ptr), ("p", self.old2new.ptr), ("i", nc)])
        self.counter_gl.fill(0)
        self.be.launch("count_num", int(nc / 512 + 1), 512, [("p", self.uniq_len.ptr), ("i", 1), ("p", self.counter_gl.ptr), ("i", nc)])
        self._L("gl_update_pos", nf, 1024, [("p", self.uniq_len.ptr), ("p", self.gl_pos.ptr), ("p", self.gl_vel.ptr),
                ("p", self.gl_pos_gen.ptr), ("p", self.gl_vel_gen.ptr), ("p", self.live.ptr), ("p", self.old2new.ptr),
                ("p", self.id_contigs.ptr), ("f", np.float32(nc - 1)), ("i", nf), ("i", id_fi), ("p", self.counter_gl.ptr),
                ("p", self.rng.ptr), ("i", 100), ("f", np.float32(0.01))])
        k2 = self.uniq_len.get()  # prefix_sum round trip (CL:79-88)
        res = np.empty_like(k2[:nc]); res[0] = 0
        if nc > 1:
            res[1:] = np.cumsum(k2[:nc - 1])
        k2[:nc] = res
        self.uniq_len.set(k2)
        return nc - 1

    def extract_uniq_mutations(self, a, b, flip_eject):  # CL:1499-1519
        self.list_uniq.fill(0)
        self.be.launch("extract_uniq_mutations", 1, 32, [("p", self.live.ptr), ("i", a), ("i", b), ("p", self.list_uniq.ptr),
                       ("p", self.valid_insert.ptr), ("p", self.n_uniq.ptr), ("i", flip_eject)])

    def pop_out_pop_in(self, a, b, mode, max_id):  # CL:1642-1778
        nf = self.nf
        self._L("pop_out_frag", nf, 1024, [("p", self.pop.ptr), ("p", self.live.ptr), ("p", self.pop_ids.ptr), ("i", a), ("i", max_id), ("i", nf)])
        max_id2 = int(self.pop_ids.get().max())  # ga.max(...).get()
        if mode == 0:
            self._L("simple_copy", nf, 1024, [("p", self.cand[0].ptr), ("p", self.pop.ptr), ("i", nf)])
        elif mode == 1:
            self._L("flip_frag", nf, 1024, [("p", self.cand[1].ptr), ("p", self.live.ptr), ("i", a), ("i", nf), ("x", 0)])
        else:
            kern = ("pop_in_frag_1", "pop_in_frag_1", "pop_in_frag_2", "pop_in_frag_2", "pop_in_frag_3", "pop_in_frag_3")[mode - 2]
            self._L(kern, nf, 1024, [("p", self.cand[mode].ptr), ("p", self.pop.ptr), ("i", a), ("i", b), ("i", max_id2),
                                     ("i", 1 if mode % 2 == 0 else -1), ("i", nf)])

    def transloc(self, a, b, max_id):  # CL:1780-1841
        nf, mode = self.nf, 0
        for up_a in range(2):
            self._L("split_contig", nf, 128, [("p", self.t1.ptr), ("p", self.live.ptr), ("p", self.t1_ids.ptr), ("i", a), ("i", up_a), ("i", max_id), ("i", nf)])
            for up_b in range(2):
                max_id1 = int(self.t1_ids.get().max())
                self._L("split_contig", nf, 128, [("p", self.t2.ptr), ("p", self.t1.ptr), ("p", self.t2_ids.ptr), ("i", b), ("i", up_b), ("i", max_id1), ("i", nf)])
                max_id2 = int(self.t2_ids.get().max())
                self._L("paste_contigs", nf, 128, [("p", self.cand[8 + mode].ptr), ("p", self.t2.ptr), ("i", a), ("i", b), ("i", max_id2), ("i", nf)])
                mode += 1

    def insert_blocks(self, a, b, max_id):  # CL:1843-1916
        nf = self.nf
        self.valid_insert.fill(-1); self.f_up.fill(-1); self.f_down.fill(-1)
        self._L("get_bounds", nf, 64, [("p", self.live.ptr), ("i", a), ("i", b), ("p", self.valid_insert.ptr), ("p", self.list_bounds.ptr),
                                        ("p", self.f_up.ptr), ("p", self.f_down.ptr), ("i", 6), ("i", nf)])
        k = 0
        for i in range(6):
            for j in (1, 0):
                lst = self.f_up if j == 1 else self.f_down
                self._L("extract_block", nf, 64, [("p", self.t1.ptr), ("p", self.live.ptr), ("p", self.t1_ids.ptr), ("i", a), ("p", lst.ptr),
                                                  ("i", i), ("i", j), ("i", max_id), ("i", nf)])
                self._L("insert_block", nf, 64, [("p", self.cand[12 + k].ptr), ("p", self.t1.ptr), ("p", self.live.ptr), ("i", a), ("i", b),
                                                 ("p", lst.ptr), ("p", self.valid_insert.ptr), ("i", k), ("i", i), ("i", j), ("i", nf)])
                k += 1

    def perform_mutations(self, a, b, max_id):
        for mode in range(8):
            self.pop_out_pop_in(a, b, mode, max_id)
        self.transloc(a, b, max_id)
        self.insert_blocks(a, b, max_id)

    def slice_sparse_mat(self, ctg1, ctg2, a, b):  # CL:1009-1069
        self.counter.fill(0)
        v = self.vect
        self._L("slice_sp_mat", self.nnz, 64, [("p", self.sp_dat.ptr), ("p", self.sp_rows.ptr), ("p", self.sp_cols.ptr), ("p", self.live.ptr),
                ("p", v[1].ptr), ("p", v[3].ptr), ("p", self.sub_rows.ptr), ("p", self.sub_cols.ptr), ("p", self.sub_dat.ptr),
                ("i", ctg1), ("i", ctg2), ("i", a), ("i", b), ("i", self.max_bounds_insert), ("p", self.counter.ptr), ("i", self.nnz)])
        n = self.n_sub_vals = int(self.counter.get()[0])
        keys, va, vb = self.sub_rows.get(), self.sub_cols.get(), self.sub_dat.get()  # sort_by_keys_zip (CL:45-55)
        idx = np.lexsort((va[:n], keys[:n])) if self.canonical_slice_order else np.argsort(keys[:n], kind="stable")
        keys[:n], va[:n], vb[:n] = keys[:n][idx], va[:n][idx], vb[:n][idx]
        self.sub_rows.set(keys); self.sub_cols.set(va); self.sub_dat.set(vb)
        self.counter.fill(0)
        self._L("prepare_sparse_call", n, 64, [("p", self.sub_rows.ptr), ("p", self.info_blocks.ptr), ("p", self.block_indptr.ptr),
                                               ("p", self.counter.ptr), ("i", n)])

    def _sub_args(self, arrays):
        return [("p", self.sub_dat.ptr), ("p", self.info_blocks.ptr), ("p", self.block_indptr.ptr), ("p", self.sub_rows.ptr),
                ("p", self.sub_cols.ptr), ("p", self.params.ptr), ("f", self.mbar)] + [("p", x.ptr) for x in arrays]

    def extract_current_sub_likelihood(self):  # CL:1156-1191
        self.cur_nz_extract.fill(0.0)
        self._L("extract_sub_likelihood", self.n_sub_vals, 64, self._sub_args(self.vect) + [("p", self.cur_nz_extract.ptr),
                ("i", self.n_sub_vals), ("i", self.ns)])
        self.cur_nz_extract.get()

    def eval_all_sub_likelihood(self):  # CL:1092-1154
        c = self.collect
        for m in range(N_TMP):  # fill_dist_all_mut CL:898-934
            self._L("fill_vect_dist", self.ns, 1024, [("p", self.s2f.ptr), ("p", self.cand[m].ptr)] + [("p", x.ptr) for x in c] + [
                    ("p", self.collector_id.ptr), ("p", self.dispatcher.ptr), ("p", self.sub_collector_id.ptr),
                    ("p", self.sub_dispatcher.ptr), ("i", self.ns), ("i", m)])
        self.vect_lik_z.fill(0); self.all_n_vals_intra.fill(0)
        self._L("eval_all_likelihood_on_zero_1st", self.ns, 1024, [("p", c[1].ptr), ("p", c[2].ptr), ("p", c[3].ptr), ("p", c[4].ptr),
                ("p", self.params.ptr), ("f", self.mbar), ("p", self.list_uniq.ptr), ("p", self.n_uniq.ptr), ("p", self.vect_lik_z.ptr),
                ("p", self.all_n_vals_intra.ptr), ("i", self.ns)])
        self.be.launch("eval_all_likelihood_on_zero_2nd", 1, 32, [("p", self.list_uniq.ptr), ("p", self.n_uniq.ptr), ("p", self.params.ptr),
                       ("p", self.vect_lik_z.ptr), ("p", self.all_n_vals_intra.ptr), ("p", self.n_pix_dev.ptr)])
        self.sub_lik_nz.fill(0.0); self.all_scores_dev.fill(0.0)
        self._L("eval_sub_likelihood", self.n_sub_vals, 64, self._sub_args(c) + [("p", self.list_uniq.ptr), ("p", self.n_uniq.ptr),
                ("p", self.sub_lik_nz.ptr), ("i", self.n_sub_vals), ("i", self.ns)])
        self.be.launch("eval_all_scores", 1, 32, [("p", self.list_uniq.ptr), ("p", self.n_uniq.ptr), ("p", self.vect_lik_z.ptr),
                       ("p", self.sub_lik_nz.ptr), ("p", self.cur_nz_extract.ptr), ("p", self.cur_nz.ptr), ("p", self.all_scores_dev.ptr)])
        return self.all_scores_dev.get()

    def test_copy_struct(self, a, b, mode, max_id):  # CL:2094-2151
        self.copy_from_gpu()
        if mode < 8:
            self.pop_out_pop_in(a, b, mode, max_id)
        elif mode < 12:
            self.transloc(a, b, max_id)
        else:
            self.insert_blocks(a, b, max_id)
        self._L("copy_struct", self.nf, 1024, [("p", self.live.ptr), ("p", self.cand[mode].ptr), ("p", self.id_contigs.ptr), ("i", self.nf)])

    def bomb(self, perm):  # CL:1925-1948
        p = DevArray(self.be, np.asarray(perm, dtype=np.int32))
        self._L("explode_genome", self.nf, 256, [("p", self.live.ptr), ("p", p.ptr), ("i", self.nf)])
        self.modify_gl_cuda_buffer(0)

    def set_valid(self, v):
        self.valid_insert.set(np.asarray(v, dtype=np.int32))

    def set_params(self, p8):
        self.params.set(np.asarray(p8, dtype=np.float32))

    def set_state(self, st13):
        for i, k in enumerate(FIELDS):
            self.live.arr[k].set(st13[i])

    def get_state(self):
        h = self.copy_from_gpu()
        return np.stack([h[k] for k in FIELDS])

    def step_sampler(self, id_frag, candidates):  # CL:1401-1465
        candidates = sorted(int(c) for c in candidates)
        n = len(candidates)
        self.fill_dist_single()
        self.eval_likelihood()
        host = self.copy_from_gpu()
        id_ctg_a = host["id_c"][id_frag]
        all_scores = np.zeros(N_TMP * n, dtype=np.float64)
        max_id = self.modify_gl_cuda_buffer(id_frag)
        flip_eject = 1
        self.n_uniq_list = []
        for i, b in enumerate(candidates):
            self.extract_uniq_mutations(id_frag, b, flip_eject)
            self.perform_mutations(id_frag, b, max_id)
            id_ctg_b = host["id_c"][b]
            self.slice_sparse_mat(id_ctg_a, id_ctg_b, id_frag, b)
            self.extract_current_sub_likelihood()
            all_scores[i * N_TMP:(i + 1) * N_TMP] = self.eval_all_sub_likelihood()
            self.n_uniq_list.append(int(np.count_nonzero(all_scores[i * N_TMP:(i + 1) * N_TMP])))
            flip_eject = 0
        self.all_scores = all_scores
        ok = np.copy(all_scores)
        ok[ok == 0] = -np.inf
        filt = ok - (ok.max() - 30)
        filt[filt < 0] = 0
        gid = int(np.argmax(filt))
        id_f_sampled = candidates[gid // N_TMP]
        op = gid % N_TMP
        self.test_copy_struct(id_frag, id_f_sampled, op, max_id)
        self.modify_gl_cuda_buffer(id_frag)
        o = all_scores[gid]
        h = self.copy_from_gpu()
        dist = dist_inter_genome(h, self.init_prev, self.init_next, self.orientable)
        return (o, dist, op, id_f_sampled, self.mean_length_contigs, self.n_contigs)

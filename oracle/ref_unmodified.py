"""TEST / BASELINE INFRASTRUCTURE ONLY: the UNMODIFIED reference ``sampler`` class on a real GPU (B-ref, route i).

The class is imported from baseline/_ref/instagraal -- a verbatim copy of /root/reference/src/instagraal that build() makes
in the container that has the reference (git-ignored, travels to the GPU box with the snapshot) -- on top of
oracle/ref_harness_gpu (a stand-in for pycuda over cuda-python, since pycuda cannot be installed here) and
oracle/_ref/ref_kernels.cubin (the reference's kernel file compiled for sm_100a).  Nothing of the product is on this path.

What this driver adds around the class is what a user of the reference does by hand: parameters are injected the way
estimate_parameters_rippe leaves them (CL:2345-2352) instead of fitted (its Python double loop over the contacts would take
hours at 1 Gb), and a scaffold can be uploaded through the class's own GPUStruct.  Every step is the class's step_sampler."""
from __future__ import annotations

import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
REF_COPY = os.path.join(_ROOT, "baseline", "_ref")
FIELDS13 = ("pos", "sub_pos", "id_c", "start_bp", "len_bp", "sub_len", "circ", "prev", "next",
            "l_cont", "sub_l_cont", "l_cont_bp", "ori")


def available():
    return (os.path.exists(os.path.join(REF_COPY, "instagraal", "cuda_lib_gl_single.py"))
            and os.path.exists(os.path.join(_HERE, "_ref", "ref_kernels.cubin"))
            and os.path.exists(os.path.join(_HERE, "_ref", "ref_signatures.json")))


def load(device=0):
    """the reference module, unmodified, with the GPU stand-in for pycuda in front of it"""
    if "pycuda" in sys.modules and "ref_harness_gpu" not in (getattr(sys.modules["pycuda"], "__file__", "") or ""):
        raise RuntimeError("another pycuda is already imported in this process")
    os.environ["IG_REF_DEVICE"] = str(int(device))
    for p in (REF_COPY, os.path.join(_HERE, "ref_harness"), os.path.join(_HERE, "ref_harness_gpu")):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REF_COPY)
    sys.path.insert(0, os.path.join(_HERE, "ref_harness"))        # empty matplotlib stand-in
    sys.path.insert(0, os.path.join(_HERE, "ref_harness_gpu"))    # pycuda stand-in (shadows ref_harness/pycuda)
    import instagraal.cuda_lib_gl_single as CL
    return CL


class UnmodifiedSampler:
    def __init__(self, level, params8, device=0):
        CL = load(device)
        import pycuda.driver as cuda
        from pycuda import _backend
        self._backend = _backend
        self.s = s = CL.sampler(*level.sampler_args())
        p = np.array([tuple(np.asarray(params8, dtype=np.float32).tolist())], dtype=s.param_simu_rippe)
        s.param_simu = p
        s.param_simu_test = p
        s.mean_value_trans = np.float32(params8[7])
        s.gpu_param_simu = cuda.mem_alloc(p.nbytes)
        s.gpu_param_simu_test = cuda.mem_alloc(p.nbytes)
        cuda.memcpy_htod(s.gpu_param_simu, s.param_simu)
        cuda.memcpy_htod(s.gpu_param_simu_test, s.param_simu_test)
        self._forced = None
        orig = s.return_neighbours

        def neighbours(id_fa, delta):   # tests hand the candidates in; the bench lets the class draw them
            return list(self._forced) if self._forced is not None else orig(id_fa, delta)
        s.return_neighbours = neighbours

    @property
    def n_launch(self):
        return self._backend.N_LAUNCH[0]

    def set_state(self, st13):
        g = self.s.gpu_vect_frags
        for i, k in enumerate(FIELDS13):
            setattr(g, k, np.ascontiguousarray(st13[i], dtype=np.int32))
        g.copy_to_gpu()

    def get_state(self):
        g = self.s.gpu_vect_frags
        g.copy_from_gpu()
        return np.stack([np.array(getattr(g, k), dtype=np.int32) for k in FIELDS13])

    def set_valid(self, v):
        self.s.gpu_list_valid_insert.set(np.asarray(v, dtype=np.int32)) if hasattr(self.s.gpu_list_valid_insert, "set") else None

    def step_sampler(self, id_frag, candidates=None, n_neighbours=5):
        self._forced = candidates
        out = self.s.step_sampler(int(id_frag), n_neighbours, np.float32(0.01))
        self._forced = None
        self.all_scores = np.asarray(self.s.all_scores, dtype=np.float64)
        self.n_sub_vals = int(self.s.n_sub_vals)
        return out

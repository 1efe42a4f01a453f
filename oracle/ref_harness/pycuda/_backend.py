import ctypes
import os

import numpy as np

_here = os.path.dirname(os.path.abspath(__file__))
_lib_path = os.environ.get(
    "IG_REF_CPU_LIB", os.path.join(_here, "..", "..", "_ref", "libref_cpu.so")
)
lib = ctypes.CDLL(os.path.abspath(_lib_path))
lib.emu_launch.restype = ctypes.c_int
lib.emu_launch.argtypes = [
    ctypes.c_void_p, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint,
    ctypes.c_int, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int, ctypes.POINTER(ctypes.c_uint32),
]

LAUNCH_LOG = []  # (kernel name) per launch, for statistics


class DeviceAllocation:
    """Host buffer posing as a device allocation; int(obj) is its address."""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self._buf = np.zeros(max(self.nbytes, 8) + 64, dtype=np.uint8)
        self.ptr = self._buf.ctypes.data

    def __int__(self):
        return self.ptr

    __index__ = __int__

    def free(self):
        pass


def as_ptr(x):
    if isinstance(x, DeviceAllocation):
        return x.ptr
    if hasattr(x, "gpudata"):
        return int(x.gpudata)
    if isinstance(x, (int, np.integer)):
        return int(x)
    raise TypeError(type(x))


_REF_CU = os.environ.get(
    "IG_REF_CU", "/root/reference/src/instagraal/kernels/kernel_sparse_adapt.cu"
)
_SIGS = None


def _signatures():
    """Parameter classes of every __global__ kernel, parsed from the reference source, so that
    arguments are marshalled POSITIONALLY like pycuda does (an np.int32 handed to a ``float``
    parameter is bit-reinterpreted, reference quirk Q1)."""
    global _SIGS
    if _SIGS is None:
        import re

        src = open(_REF_CU).read()
        src = re.sub(r"//[^\n]*", "", src)
        _SIGS = {}
        for m in re.finditer(r"__global__\s+void\s+(\w+)\s*\(([^)]*)\)", src):
            kinds = []
            for prm in m.group(2).split(","):
                prm = prm.strip()
                if not prm:
                    continue
                if "*" in prm:
                    kinds.append("p")
                elif re.search(r"\bfloat2\b", prm):
                    kinds.append("x")
                elif re.search(r"\bfloat\b", prm):
                    kinds.append("f")
                elif re.search(r"\bdouble\b", prm):
                    kinds.append("d")
                elif "long" in prm:
                    kinds.append("q")
                else:
                    kinds.append("i")
            _SIGS[m.group(1)] = kinds
    return _SIGS


def launch(name, fn_addr, args, block, grid):
    kinds = _signatures()[name]
    ints, fps = [], []
    for a, k in zip(args, kinds):
        if k == "p":
            ints.append(as_ptr(a))
        elif k == "f":
            if isinstance(a, np.float32):
                fps.append(int(a.view(np.uint32)))
            elif isinstance(a, (np.int32, np.uint32)):
                fps.append(int(np.int32(a).view(np.uint32)))  # bit reinterpretation (Q1)
            else:
                raise TypeError("bad arg %r for float param of %s" % (type(a), name))
        elif k in ("i", "q"):
            if isinstance(a, np.float32):
                ints.append(int(a.view(np.uint32)))
            else:
                ints.append(int(a) & 0xFFFFFFFFFFFFFFFF)
        else:
            raise TypeError("unsupported param kind %s in %s" % (k, name))
    ia = (ctypes.c_uint64 * max(len(ints), 1))(*ints)
    fa = (ctypes.c_uint32 * max(len(fps), 1))(*fps)
    bx, by, bz = (list(block) + [1, 1, 1])[:3]
    gx, gy = (list(grid) + [1, 1])[:2]
    LAUNCH_LOG.append(name)
    rc = lib.emu_launch(fn_addr, int(gx), int(gy), int(bx), int(by), int(bz), len(ints), ia, len(fps), fa)
    if rc != 0:
        raise RuntimeError("emu_launch(%s) failed rc=%d" % (name, rc))

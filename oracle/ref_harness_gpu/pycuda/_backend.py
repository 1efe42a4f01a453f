import ctypes
import json
import os

import numpy as np
from cuda.bindings import driver as drv

_here = os.path.dirname(os.path.abspath(__file__))
_ref = os.path.abspath(os.path.join(_here, "..", "..", "_ref"))


def ck(res):
    if int(res[0]) != 0:
        raise RuntimeError("CUDA driver error %s" % res[0])
    return res[1] if len(res) == 2 else (res[1:] if len(res) > 2 else None)


ck(drv.cuInit(0))
_dev = ck(drv.cuDeviceGet(int(os.environ.get("IG_REF_DEVICE", "0"))))
ctx = ck(drv.cuDevicePrimaryCtxRetain(_dev))
ck(drv.cuCtxSetCurrent(ctx))
_module = ck(drv.cuModuleLoadData(open(os.path.join(_ref, "ref_kernels.cubin"), "rb").read()))
SIGS = json.load(open(os.path.join(_ref, "ref_signatures.json")))   # kernel -> parameter classes (oracle/make_signatures.py)
N_LAUNCH = [0]


class DeviceAllocation:
    def __init__(self, nbytes):
        self.nbytes = max(int(nbytes), 8)
        self.ptr = int(ck(drv.cuMemAlloc(self.nbytes)))

    def __int__(self):
        return self.ptr

    __index__ = __int__

    def free(self):
        if self.ptr:
            drv.cuMemFree(self.ptr)
            self.ptr = 0


def as_ptr(x):
    if isinstance(x, DeviceAllocation):
        return x.ptr
    if hasattr(x, "gpudata"):
        return int(x.gpudata)
    if isinstance(x, (int, np.integer)):
        return int(x)
    raise TypeError(type(x))


def function(name):
    return ck(drv.cuModuleGetFunction(_module, name.encode()))


_CT = {"p": ctypes.c_void_p, "f": ctypes.c_float, "i": ctypes.c_int, "q": ctypes.c_ulonglong, "d": ctypes.c_double, "x": ctypes.c_double}


def launch(name, fn, args, block, grid):
    """arguments are marshalled POSITIONALLY like pycuda does: 4-byte scalars keep their bits whatever the parameter's
    type (an np.int32 handed to a ``float`` parameter is reinterpreted, reference quirk Q1); parameters the caller
    leaves out (flip_frag's unused trailing float2) are zero."""
    kinds = SIGS[name]
    vals, types = [], []
    for i, k in enumerate(kinds):
        a = args[i] if i < len(args) else 0
        if k == "p":
            vals.append(as_ptr(a)); types.append(ctypes.c_void_p)
        elif k in ("f", "i"):
            if isinstance(a, np.float32):
                vals.append(float(a)); types.append(ctypes.c_float)
            elif isinstance(a, (np.float64, float)) and k == "f":
                vals.append(float(a)); types.append(ctypes.c_float)
            else:
                vals.append(int(np.int32(a))); types.append(ctypes.c_int)
        elif k == "d":
            vals.append(float(a)); types.append(ctypes.c_double)
        elif k == "x":
            vals.append(0.0); types.append(ctypes.c_double)
        else:
            vals.append(int(a)); types.append(ctypes.c_ulonglong)
    bx, by, bz = (list(block) + [1, 1, 1])[:3]
    gx, gy = (list(grid) + [1, 1])[:2]
    ck(drv.cuLaunchKernel(fn, int(gx), int(gy), 1, int(bx), int(by), int(bz), 0, 0, (tuple(vals), tuple(types)), 0))
    N_LAUNCH[0] += 1

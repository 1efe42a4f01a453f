import numpy as np

from ._backend import DeviceAllocation, as_ptr, ck, drv


def mem_alloc(nbytes):
    return DeviceAllocation(nbytes)


def mem_alloc_like(ary):
    return DeviceAllocation(ary.nbytes)


def memcpy_htod(dest, src):
    if isinstance(src, (bytes, bytearray)):
        a = np.frombuffer(bytes(src), dtype=np.uint8)
    else:
        a = np.ascontiguousarray(src)
    if a.nbytes:
        ck(drv.cuMemcpyHtoD(as_ptr(dest), a.ctypes.data, a.nbytes))


def memcpy_dtoh(dest, src):
    assert isinstance(dest, np.ndarray)
    if dest.nbytes:
        ck(drv.cuMemcpyDtoH(dest.ctypes.data, as_ptr(src), dest.nbytes))


def to_device(bf):
    a = np.frombuffer(bytes(bf), dtype=np.uint8) if isinstance(bf, (bytes, bytearray)) else np.ascontiguousarray(bf)
    d = DeviceAllocation(a.nbytes)
    memcpy_htod(d, a)
    return d


def mem_get_info():
    free, total = ck(drv.cuMemGetInfo())
    return int(free), int(total)


class Event:
    """the reference brackets every launch with start.record() ... end.record(); end.synchronize()"""

    def record(self, stream=None):
        return self

    def synchronize(self):
        ck(drv.cuCtxSynchronize())
        return self

    def time_till(self, other):
        return 0.0

    def time_since(self, other):
        return 0.0


class _Ctx:
    def synchronize(self):
        ck(drv.cuCtxSynchronize())

    def detach(self):
        pass


class Context:
    _c = _Ctx()

    @staticmethod
    def get_current():
        return Context._c

    @staticmethod
    def synchronize():
        ck(drv.cuCtxSynchronize())

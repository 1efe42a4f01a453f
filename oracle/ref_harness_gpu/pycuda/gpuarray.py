import numpy as np

from ._backend import DeviceAllocation, ck, drv


class GPUArray:
    def __init__(self, shape, dtype, data=None):
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape)) if len(self.shape) else 1
        self.nbytes = self.size * self.dtype.itemsize
        self._alloc = DeviceAllocation(self.nbytes)
        if data is not None:
            self.set(data)

    @property
    def gpudata(self):
        return self._alloc.ptr

    def set(self, ary):
        a = np.ascontiguousarray(ary, dtype=self.dtype)
        if a.nbytes:
            ck(drv.cuMemcpyHtoD(self._alloc.ptr, a.ctypes.data, a.nbytes))
        return self

    def get(self, ary=None):
        out = np.empty(self.shape, dtype=self.dtype) if ary is None else ary
        if out.nbytes:
            ck(drv.cuMemcpyDtoH(out.ctypes.data, self._alloc.ptr, out.nbytes))
        return out

    def fill(self, v):
        return self.set(np.full(self.shape, v, dtype=self.dtype))

    def __len__(self):
        return self.shape[0] if self.shape else 1


def to_gpu(ary):
    a = np.ascontiguousarray(ary)
    return GPUArray(a.shape, a.dtype, a)


def zeros(shape, dtype=np.float32):
    if not isinstance(shape, tuple):
        shape = (int(shape),)
    shape = tuple(int(s) for s in np.ravel(shape))
    return GPUArray(shape, dtype, np.zeros(shape, dtype=dtype))


def zeros_like(other):
    return zeros(other.shape, other.dtype)


def max(a, stream=None):
    return to_gpu(np.array(a.get().max()))


def sum(a, dtype=None, stream=None):
    return to_gpu(np.array(a.get().sum(dtype=dtype)))

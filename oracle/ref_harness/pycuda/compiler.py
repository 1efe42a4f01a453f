import ctypes

from . import _backend


class _Function:
    def __init__(self, name):
        self.name = name
        self.addr = ctypes.cast(getattr(_backend.lib, name), ctypes.c_void_p).value

    def __call__(self, *args, block=(1, 1, 1), grid=(1, 1), shared=0, **kw):
        _backend.launch(self.name, self.addr, args, block, grid)


class SourceModule:
    """Ignores the (macro-substituted) source text: the same file was compiled in place with the
    same five macro values passed as -D (oracle/Makefile)."""

    def __init__(self, source, **kw):
        self.source = source

    def get_function(self, name):
        return _Function(name)

from . import _backend


class _Function:
    def __init__(self, name):
        self.name = name
        self.fn = _backend.function(name)

    def __call__(self, *args, block=(1, 1, 1), grid=(1, 1), shared=0, **kw):
        _backend.launch(self.name, self.fn, args, block, grid)


class SourceModule:
    """Ignores the (macro-substituted) source text: the same file was compiled for sm_100a with the same five macro
    values passed as -D (oracle/Makefile -> oracle/_ref/ref_kernels.cubin)."""

    def __init__(self, source, **kw):
        self.source = source

    def get_function(self, name):
        return _Function(name)

def __getattr__(name):
    raise RuntimeError("matplotlib stub: plotting is not available in the reference harness")

"""The C-ABI library loads and exports every symbol include/instagraal_b200.h declares; without a
GPU the product fails loudly instead of falling back to a CPU path."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "instagraal_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ig_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported(built):
    from instagraal_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s
    assert set(syms) == set(_lib.EXPORTS)


def test_no_cpu_fallback_without_gpu(built):
    from instagraal_b200 import _lib
    L = _lib.lib()
    if L.ig_device_count() > 0:
        pytest.skip("a GPU is present")
    import numpy as np
    from instagraal_b200.cuda_lib_gl_single import sampler
    from instagraal_b200.synth import WORKLOADS, make_level
    level = make_level(WORKLOADS["micro"])
    with pytest.raises(RuntimeError, match="no CUDA device|ig_create"):
        sampler(*level.sampler_args())
    assert np.int32(0) == 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "instagraal_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f

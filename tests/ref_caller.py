"""TEST INFRASTRUCTURE: run the reference's own caller of the hot path -- the UNMODIFIED class `simulation` of
instagraal/simu_single.py (the verbatim copy under baseline/_ref/, or /root/reference/src where it exists) -- with its two
collaborators chosen by the test:
    instagraal.pyramid_sparse        ->  the reference module (over an in-memory h5py stand-in)   or   our pyramid_build + pyramid_load
    instagraal.cuda_lib_gl_single    ->  a recorder of the 29 constructor arguments               or   our `sampler` facade
`simulation.__init__` is exactly what a user's `instagraal` run executes before the MCMC loop (IG:556-566): build_and_filter,
two levels, sequences per bin, sub-fragment tables, the sampler constructor, estimate_parameters_rippe.
"""
import importlib
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_src():
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"):
        if os.path.isdir(os.path.join(p, "instagraal")):
            return p
    return None


# ---- h5py stand-in (h5py is not installed here): files live in a per-process registry keyed by path -------------------------
class _Group(dict):
    def create_dataset(self, name, shape, dtype):
        self[name] = np.zeros(shape, dtype=np.int32)
        return self[name]


class _File(dict):
    _registry = {}

    def __new__(cls, path, mode="a"):
        path = os.path.abspath(str(path))
        if path not in cls._registry:
            obj = dict.__new__(cls)
            obj.attrs = {}
            cls._registry[path] = obj
            open(path, "a").close()
        return cls._registry[path]

    def __init__(self, path, mode="a"):
        pass

    def create_group(self, name):
        self[name] = _Group()
        return self[name]

    def close(self):
        pass


def fake_h5py():
    m = types.ModuleType("h5py")
    m.File = _File
    return m


class _Cmap:
    """plt.cm.prism as simu_single.py:233-239 uses it (colours of the dead OpenGL viewer; never reach the sampler)"""
    N = 256

    def __call__(self, i):
        return (0.0, 0.0, 0.0, 1.0)


class _NoPlot(types.ModuleType):
    cm = types.SimpleNamespace(prism=_Cmap(), gist_ncar=_Cmap())

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None

    @staticmethod
    def subplots(*a, **k):   # what this repo's display_current_matrix asks for first: no matplotlib here -> its PGM branch
        raise ImportError("matplotlib stand-in: no drawing")


def _purge():
    for k in [k for k in sys.modules if k == "instagraal" or k.startswith("instagraal.")]:
        del sys.modules[k]


class RecordingSampler:
    """stands where `sampler` stands in simu_single.py:120-153 and keeps what it was given"""
    last = None

    def __init__(self, *args):
        assert len(args) == 29, len(args)
        self.args = args
        soa = args[1]
        self.gpu_vect_frags = types.SimpleNamespace(start_bp=np.asarray(soa["start_bp"]), l_cont_bp=np.asarray(soa["l_cont_bp"]))
        self.rippe_call = None
        RecordingSampler.last = self

    def estimate_parameters_rippe(self, max_dist_kb, size_bin_kb, display_graph):
        self.rippe_call = (max_dist_kb, size_bin_kb, display_graph)

    def simulate_rippe_contacts(self, *a):
        raise AssertionError("is_simu is False on the live path")


def our_pyramid_module():
    """instagraal.pyramid_sparse as a maintainer would alias it: build side + load side of this repo under one name"""
    from instagraal_b200 import pyramid_build, pyramid_load
    m = types.ModuleType("instagraal.pyramid_sparse")
    for src in (pyramid_build, pyramid_load):
        for k, v in vars(src).items():
            if not k.startswith("_"):
                setattr(m, k, v)
    return m


def load_simulation(pyr="reference", sampler_cls=RecordingSampler):
    """returns the reference's `simulation` class wired to the chosen collaborators (fresh import every time)"""
    src = reference_src()
    if src is None:
        raise RuntimeError("no copy of the reference package (baseline/_ref or /root/reference)")
    _purge()
    if src not in sys.path:
        sys.path.insert(0, src)
    try:   # a real matplotlib, if there is one; the raising stub of oracle/ref_harness (which other harnesses put on sys.path) is not
        import matplotlib
        real = hasattr(matplotlib, "__version__")
        if real:
            import matplotlib.pyplot  # noqa: F401
    except Exception:
        real = False
    if not real:
        mp = types.ModuleType("matplotlib")
        mp.pyplot = _NoPlot("matplotlib.pyplot")
        mp.use = lambda *a, **k: None
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mp, mp.pyplot
    pkg = importlib.import_module("instagraal")
    if pyr == "reference":
        if "h5py" not in sys.modules:
            try:
                import h5py  # noqa: F401
            except ImportError:
                sys.modules["h5py"] = fake_h5py()
        pyr_mod = importlib.import_module("instagraal.pyramid_sparse")
        pyr_mod.plt = _NoPlot("plt")   # remove_problematic_fragments' two diagnostic PDFs
    else:
        pyr_mod = our_pyramid_module()
        sys.modules["instagraal.pyramid_sparse"] = pyr_mod
    pkg.pyramid_sparse = pyr_mod
    cl = types.ModuleType("instagraal.cuda_lib_gl_single")
    cl.sampler = sampler_cls
    sys.modules["instagraal.cuda_lib_gl_single"] = cl
    pkg.cuda_lib_gl_single = cl
    simu = importlib.import_module("instagraal.simu_single")
    assert simu.sampler_lib is sampler_cls and simu.pyr is pyr_mod
    return simu.simulation


def write_dataset(folder, seed=5, n_frags=(60, 90, 130, 40, 200, 7)):
    """an `instagraal-pre` output folder + the FASTA it was made from (random sequences of the contigs' lengths)"""
    sys.path.insert(0, ROOT)
    from oracle.make_pyramid_golden import write_input
    write_input(folder, seed=seed, n_frags=n_frags)
    rng = np.random.RandomState(seed + 1)
    fasta = os.path.join(folder, "genome.fa")
    with open(os.path.join(folder, "info_contigs.txt")) as h, open(fasta, "w") as out:
        h.readline()
        for ln in h:
            name, length = ln.split("\t")[:2]
            s = "".join(rng.choice(list("ACGT"), int(length)))
            out.write(">%s\n" % name)
            for i in range(0, len(s), 80):
                out.write(s[i:i + 80] + "\n")
        out.write("\n")   # (load_reference_sequence drops the file's last line, PS:1649: keep the last record whole)
    return fasta


def load_run_instagraal(sampler_cls):
    """the reference's whole program, `instagraal.run_instagraal` (IG:502-600: what the `instagraal` command line calls),
    unmodified, wired to this repo's pyramid modules and `sampler_cls`.  Stubbed besides: `pycuda.autoinit` (a context for
    pycuda, which nothing here uses) and the post-run report `assembly_stats` (Biopython, not installed)."""
    load_simulation("ours", sampler_cls)
    for name in ("pycuda", "pycuda.autoinit"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    st = types.ModuleType("instagraal.assembly_stats")
    st.print_assembly_stats = lambda *a, **k: None
    sys.modules["instagraal.assembly_stats"] = st
    sys.modules["instagraal"].assembly_stats = st
    ig = importlib.import_module("instagraal.instagraal")
    return ig.run_instagraal

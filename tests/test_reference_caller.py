"""Drop-in proof on the CALLER's side: the reference's own, unmodified `simulation` class (simu_single.py -- what every
`instagraal` run executes before the MCMC loop) is run twice on the same pre-processing output,
  (1) with the reference's pyramid_sparse module (h5py replaced by an in-memory stand-in), and
  (2) with this repo's pyramid_build + pyramid_load aliased in its place,
and must hand the sampler constructor the same 29 arguments, bit for bit.  On the GPU box the same class then constructs THIS
repo's `sampler`, estimates the p(s) parameters, and runs the first steps of full_em (IG:204-284) and the FASTA export."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import ref_caller as RC

needs_ref = pytest.mark.skipif(RC.reference_src() is None, reason="no copy of the reference package (baseline/_ref is made by build())")

ARG_NAMES = ["use_rippe", "S_o_A_frags", "collector_id_repeats", "frag_dispatcher", "id_frag_duplicated", "id_frags_blacklisted", "n_frags",
             "n_new_frags", "init_n_sub_frags", "n_new_sub_frags", "np_rep_sub_frags_id", "sub_sampled_sparse_matrix", "np_sub_frags_len_bp",
             "np_sub_frags_id", "np_sub_frags_accu", "np_sub_frags_2_frags", "mean_squared_frags_per_bin", "norm_vect_accu", "sub_candidates_dup",
             "sub_candidates_output_data", "S_o_A_sub_frags", "sub_collector_id_repeats", "sub_frag_dispatcher", "sparse_matrix",
             "mean_value_trans", "n_iterations", "is_simu", "vel", "pos"]


@pytest.fixture(autouse=True)
def _restore_modules():
    """the harness plants stand-ins (h5py, matplotlib, pycuda, instagraal.*) in sys.modules: take them out again"""
    import sys
    before = dict(sys.modules)
    yield
    for k in list(sys.modules):
        if k.split(".")[0] in ("h5py", "matplotlib", "pycuda", "instagraal"):
            if k in before:
                sys.modules[k] = before[k]
            else:
                del sys.modules[k]


def same(a, b, path):
    if sp.issparse(a) or sp.issparse(b):
        assert sp.issparse(a) and sp.issparse(b) and a.shape == b.shape and a.dtype == b.dtype and a.format == b.format, path
        for k in ("data", "indices", "indptr"):
            assert np.array_equal(getattr(a, k), getattr(b, k)), (path, k)
    elif isinstance(a, dict):
        assert isinstance(b, dict) and list(a.keys()) == list(b.keys()), path
        for k in a:
            same(a[k], b[k], path + "[%r]" % (k,))
    elif isinstance(a, np.ndarray):
        assert isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape, (path, a.dtype, getattr(b, "dtype", None))
        assert a.tobytes() == b.tobytes(), path
    elif isinstance(a, (list, tuple)):
        assert type(a) is type(b) and len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            same(x, y, path + "[%d]" % i)
    else:
        assert type(a) is type(b), (path, type(a), type(b))
        assert (a == b) or (a != a and b != b), (path, a, b)


def _numpy_bin_contacts(*a, **k):
    from test_pyramid_build import _numpy_bin_contacts as f
    return f(*a, **k)


@needs_ref
def test_simulation_hands_the_sampler_the_same_arguments(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)          # the reference's logger and its diagnostic plots write into the working directory
    base = str(tmp_path / "pre")
    fasta = RC.write_dataset(base)
    got = {}
    for which in ("reference", "ours"):
        if which == "ours":
            from instagraal_b200 import pyramid_build as pb
            monkeypatch.setattr(pb, "bin_contacts", _numpy_bin_contacts)   # host logic only: no GPU in this test
            from instagraal_b200 import pyramid_load as pl
            monkeypatch.setattr(pl._TextBackedData, "__getitem__", _text_backed_getitem_numpy)
        simulation = RC.load_simulation(which, RC.RecordingSampler)
        np.random.seed(4)
        sim = simulation("ds", base, fasta, 2, 10, False, True, thresh_factor=1, output_folder=str(tmp_path / ("out_" + which)))
        rec = RC.RecordingSampler.last
        got[which] = (rec.args, rec.rippe_call, sim)
    a_ref, call_ref, sim_ref = got["reference"]
    a_our, call_our, sim_our = got["ours"]
    for name, x, y in zip(ARG_NAMES, a_ref, a_our):
        same(x, y, name)
    same(call_ref, call_our, "estimate_parameters_rippe args")
    assert sim_ref.n_frags == sim_our.n_frags > 20 and sim_ref.init_n_sub_frags == sim_our.init_n_sub_frags
    same(sim_ref.level.list_seq, sim_our.level.list_seq, "level.list_seq")
    for k in ("new_fasta", "info_frags"):
        assert os.path.basename(getattr(sim_ref, k)) == os.path.basename(getattr(sim_our, k))


def _text_backed_getitem_numpy(self, key):
    """_TextBackedData.__getitem__ with the GPU binning replaced by the test-local NumPy statement (CPU run of the host logic)"""
    if key not in self.cache:
        from instagraal_b200 import pyramid_build as pb
        lvl = int(key)
        n = int(self.owner.spec_level[key]["frag_columns"]["index"].size)
        arr = pb.fill_sparse_pyramid_level(None, lvl, os.path.join(self.owner.spec_level[key]["level_folder"], "%d_abs_frag_contacts.txt" % lvl), n)
        self.cache[key] = {"data": arr, "nfrags": np.array([[n]], dtype=np.int32)}
    return self.cache[key]


@needs_ref
@pytest.mark.gpu
def test_reference_simulation_drives_this_sampler_on_the_gpu(built, tmp_path, monkeypatch):
    """the import swap of INTEGRATION.md section 1, executed: reference `simulation` + this repo's pyramid modules + this repo's
    `sampler`; then the head of full_em (IG:204-284): bomb, a sweep of step_sampler with step_nuisance_parameters, FASTA export"""
    monkeypatch.chdir(tmp_path)
    base = str(tmp_path / "pre")
    fasta = RC.write_dataset(base)
    from instagraal_b200.cuda_lib_gl_single import sampler
    simulation = RC.load_simulation("ours", sampler)
    np.random.seed(4)
    sim = simulation("ds", base, fasta, 2, 10, False, True, thresh_factor=1, output_folder=str(tmp_path / "out"))
    s = sim.sampler
    assert isinstance(s, sampler) and int(s.n_new_frags) == sim.n_frags
    kuhn, lm, c1, slope, d, d_max, fact, d_nuc = s.param_simu[0]
    assert slope < 0 and d_max > 0 and fact > 0
    s.bomb_the_genome()
    frs = np.arange(sim.n_frags, dtype=np.int32)
    np.random.shuffle(frs)
    n_contigs = []
    for j, f in enumerate(frs):
        o, dist, op, id_f, mean_len, nc = s.step_sampler(int(f), 5, np.float32(0.01))
        assert np.isfinite(o) and 0 <= op < 24 and 0 <= id_f < sim.n_frags
        n_contigs.append(int(nc))
        fact_, d_, d_max_, d_nuc_, slope_, lik, success, y_rippe = s.step_nuisance_parameters(np.float32(0.01), 0, j)
        assert np.isfinite(lik) and success in (0, 1)
    assert n_contigs[-1] < sim.n_frags          # the exploded genome has been merging
    s.gpu_vect_frags.copy_from_gpu()
    sim.export_new_fasta()
    txt = open(sim.new_fasta).read()
    assert txt.startswith(">3C-assembly-contig_")
    n_bases = sum(len(l) for l in txt.splitlines() if not l.startswith(">"))
    total = sum(len(q) for q in sim.level.list_seq)
    assert total - 2 * txt.count(">") <= n_bases <= total      # (the writer's 1-character last-line quirk can drop a base per contig)
    assert open(sim.info_frags).read().count("init_contig\tid_frag") == txt.count(">")
    s.free_gpu()


@needs_ref
@pytest.mark.gpu
def test_the_reference_program_runs_end_to_end_on_this_repo(built, tmp_path, monkeypatch):
    """`instagraal.run_instagraal` itself (IG:502-600, the body of the `instagraal` command), unmodified: pyramid build with
    filtering, load, sampler construction, p(s) fit, bomb, 6 cycles of full_em (the nuisance step from cycle 5 on), per-cycle
    FASTA / info_frags / thumbnails / behaviour files -- every collaborator below `simulation` is this repo's"""
    monkeypatch.chdir(tmp_path)
    base = str(tmp_path / "pre")
    fasta = RC.write_dataset(base)
    from instagraal_b200.cuda_lib_gl_single import sampler
    run = RC.load_run_instagraal(sampler)
    np.random.seed(11)
    run(base, fasta, str(tmp_path / "out"), level=2, cycles=6, coverage_std=1, neighborhood=5, device=0, bomb=True, save_matrix=True)
    res = os.path.join(str(tmp_path / "out"), "pre", "test_mcmc_2")
    n_frags = sum(1 for _ in open(os.path.join(res, "save_simu_step_5.txt")))
    assert n_frags > 20
    lik = [float(x) for x in open(os.path.join(res, "list_likelihood.txt"))]
    assert len(lik) == 6 * n_frags and np.all(np.isfinite(lik))
    assert lik[-1] > lik[0]                                           # the assembly improves from the exploded start
    n_contigs = [int(x) for x in open(os.path.join(res, "list_n_contigs.txt"))]
    assert n_contigs[-1] < n_frags // 2
    assert len(open(os.path.join(res, "list_fact.txt")).readlines()) == n_frags          # nuisance steps: cycle 5 only (j > 4)
    assert set(open(os.path.join(res, "list_success.txt")).read().split()) <= {"0", "1"}
    muts = open(os.path.join(res, "list_mutations.txt")).read().splitlines()
    assert muts[0] == "id_fA\tid_fB\tid_mutation" and len(muts) == 1 + 6 * n_frags
    fa = open(os.path.join(res, "genome.fasta")).read()
    assert fa.count(">3C-assembly-contig_") == n_contigs[-1]
    assert open(os.path.join(res, "info_frags.txt")).read().count(">3C-assembly|contig_") == n_contigs[-1]
    for j in range(6):
        assert os.path.getsize(os.path.join(res, "matrix_cycle_%d.png" % j)) > 1000   # (PGM bytes under the reference's file name: no matplotlib here)
    _structural_pins_of_the_reference_gpu_tests(res, n_cycles=6, n_frags=n_frags)


def _structural_pins_of_the_reference_gpu_tests(res, n_cycles, n_frags):
    """the assertions of the reference's own tests/test_instagraal_gpu.py:128-330 (its only pins of this path: structural),
    on this run's output folder"""
    import glob
    import math
    n_iters = n_cycles * n_frags
    for fname in ("genome.fasta", "info_frags.txt", "list_likelihood.txt", "list_n_contigs.txt", "list_mean_len.txt", "list_dist_init_genome.txt",
                  "list_mutations.txt", "save_simu_step_0.txt", "save_simu_step_%d.txt" % (n_cycles - 1), "matrix_cycle_0.png",
                  "matrix_cycle_%d.png" % (n_cycles - 1)):
        assert os.path.exists(os.path.join(res, fname)), fname
    lines = open(os.path.join(res, "genome.fasta")).read().splitlines()
    headers = [l for l in lines if l.startswith(">")]
    assert len(headers) >= 1
    for h in headers:
        assert h[1:].startswith("3C-assembly-contig_") and h[1:].split("3C-assembly-contig_")[1].isdigit()
    assert all(set(l) <= set("ACGTNacgtn") for l in lines if not l.startswith(">"))
    current, started = None, False
    for l in lines:
        if l.startswith(">"):
            assert current is None or started, "empty sequence for %s" % current
            current, started = l, False
        elif l.strip():
            started = True
    assert started
    blocks = [b.strip() for b in open(os.path.join(res, "info_frags.txt")).read().split(">") if b.strip()]
    assert len(blocks) >= 1
    for b in blocks:
        bl = b.splitlines()
        assert len(bl) >= 2 and set(bl[1].split()) == {"init_contig", "id_frag", "orientation", "start", "end"}
        for row in bl[2:]:
            f = row.split()
            assert len(f) == 5 and int(f[2]) in (1, -1)
    assert len(glob.glob(os.path.join(res, "save_simu_step_*.txt"))) == n_cycles
    for i in range(n_cycles):
        rows = open(os.path.join(res, "save_simu_step_%d.txt" % i)).read().splitlines()
        assert len(rows) == n_frags
    for l in open(os.path.join(res, "save_simu_step_0.txt")).read().splitlines():
        f = l.split()
        assert len(f) == 4 and int(f[3]) in (1, -1) and all(int(x) == int(x) for x in f)
    for name in ("list_likelihood.txt", "list_n_contigs.txt", "list_mean_len.txt", "list_dist_init_genome.txt"):
        vals = [v for v in open(os.path.join(res, name)).read().splitlines() if v.strip()]
        assert len(vals) == n_iters, name
        if name == "list_likelihood.txt":
            assert all(math.isfinite(float(v)) for v in vals)
        if name == "list_n_contigs.txt":
            assert all(int(v) > 0 for v in vals)
    mut = [l.split("\t") for l in open(os.path.join(res, "list_mutations.txt")).read().splitlines()]
    assert mut[0] == ["id_fA", "id_fB", "id_mutation"] and len(mut) - 1 == n_iters
    body = np.array(mut[1:], dtype=np.int64)
    assert body[:, :2].min() >= 0 and body[:, :2].max() <= n_frags - 1 and body[:, 2].min() >= 0
    assert len(glob.glob(os.path.join(res, "matrix_cycle_*.png"))) == n_cycles

"""Replica chains on the GPU: several chains per device in one call (ig_clone + ig_run_cycles_device_multi) and the
all-gather of their likelihoods / scaffolds inside the library (ig_allgather_best; NCCL when there are >= 2 GPUs)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from instagraal_b200.synth import WORKLOADS, make_level

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P8 = np.array([2.2354, 1.4933294, 0.06928191, -0.9384134, 2.0, 386.88467, 65.71848, 0.01698581], dtype=np.float32)


def _orders(nf, n_chains, n_cycles):
    rng = np.random.RandomState(77)
    return np.stack([np.concatenate([rng.permutation(nf) for _ in range(n_cycles)]) for _ in range(n_chains)]).astype(np.int32)


def test_chains_sharing_a_level_equal_separate_handles(built):
    """4 chains of one level driven together == the same 4 chains each on its own handle, record for record, and the
    final scaffolds are identical; the all-gather (single rank) returns every chain's likelihood and scaffold."""
    from instagraal_b200.replicas import ReplicaSet, best_chain
    from test_gpu_parity import make_sampler
    level = make_level(WORKLOADS["toy"])
    n, n_cyc = 4, 2
    frags = _orders(level.n_frags, n, n_cyc)
    seeds = np.array([11, 12, 13, 14], dtype=np.uint64)
    # reference: separate handles, one after the other
    want, want_state = [], []
    for i in range(n):
        s = make_sampler(level)
        s.set_gpu_share(4)   # what ReplicaSet gives each of 4 chains (the share groups the partial sums, compare like with like)
        s.set_param_simu(P8)
        np.random.seed(100 + i)
        s.bomb_the_genome()
        want.append(s.run_cycle_device(frags[i], 5, seed=int(seeds[i]), cycle=3))
        want_state.append(s._get_state())
        s.free_gpu()
    first = make_sampler(level)
    first.set_param_simu(P8)
    rs = ReplicaSet(first, n, seeds=seeds)
    rs.bomb(seed0=100)
    got = rs.run_cycle(frags, 5, cycle=3)
    for i in range(n):
        for key in ("likelihood", "dist", "op_sampled", "id_f_sampled", "n_contigs", "n_proposals"):
            assert np.array_equal(got[i][key], want[i][key]), (i, key)
        assert np.array_equal(rs.chains[i]._get_state(), want_state[i]), i
    best, lik, nc = rs.allgather()
    assert rs.n_gathers == 1 and rs.gather_ms > 0
    assert np.array_equal(lik, np.array([got[i][-1]["likelihood"] for i in range(n)]))
    assert np.array_equal(nc, np.array([got[i][-1]["n_contigs"] for i in range(n)]))
    assert best == best_chain(lik)
    for i in range(n):
        assert np.array_equal(rs.gathered_state(i), want_state[i]), i
    rs.free()


def test_nccl_allgather_two_gpus(built, tmp_path):
    """world_size 2, one process per GPU, 2 chains each: the NCCL all-gather inside the library delivers the same table
    and the same best scaffold to both ranks."""
    from instagraal_b200 import _lib as L
    if L.lib().ig_device_count() < 2:
        pytest.skip("needs two GPUs")
    worker = os.path.join(ROOT, "tests", "_nccl_worker.py")
    idfile = str(tmp_path / "nccl_id.bin")
    ps = [subprocess.Popen([sys.executable, worker, str(r), "2", idfile], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
          for r in range(2)]
    res = []
    for p in ps:
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out[-3000:]
        res.append(json.loads([ln for ln in out.splitlines() if ln.startswith("RESULT ")][-1][7:]))
    a, b = sorted(res, key=lambda r: r["rank"])
    assert a["lik"] == b["lik"] and a["nc"] == b["nc"] and a["best"] == b["best"]
    assert len(a["lik"]) == 4
    assert a["best_state_crc"] == b["best_state_crc"]
    # rank-major layout: rank r's own chains sit at [2r, 2r + 2)
    assert a["own"] == a["lik"][0:2] and b["own"] == b["lik"][2:4]

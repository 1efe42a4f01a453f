"""Host-side replica logic, world_size 2, on CPU.  The collective of the product is ncclAllGather inside the library
(ig_allgather_best; GPU tests in test_gpu_replicas.py); what can be checked without a GPU is the launcher-side contract:
  * every rank derives the same decision (best chain) from the same gathered table (best_chain),
  * the chain -> (rank, local index) layout is rank-major with n_local chains per rank,
  * a 128-byte id created by one rank reaches the others unchanged (here over torch.distributed/gloo, which is what
    bench.py uses for that one broadcast).
"""
import os
import socket

import numpy as np

from instagraal_b200.replicas import best_chain


def test_best_chain_is_deterministic():
    assert best_chain([-5.0, -3.0, -3.0, -9.0]) == 1
    assert best_chain([-1.0]) == 0
    assert best_chain(np.array([-2.0, -2.0])) == 0


def test_launcher_contract_world_size_2_gloo():
    import json
    import subprocess
    import sys
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    worker = os.path.join(os.path.dirname(__file__), "_replica_worker.py")
    ps = [subprocess.Popen([sys.executable, worker, str(r), "2", str(port)], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True) for r in range(2)]
    res = []
    for p in ps:
        out, _ = p.communicate(timeout=300)
        assert p.returncode == 0, out[-2000:]
        line = [ln for ln in out.splitlines() if ln.startswith("RESULT ")][-1]
        res.append(json.loads(line[7:]))
    assert len(res) == 2
    for rank, best, liks, id_sum in sorted(res):
        assert best == 5                       # chain 1 of rank 1 (rank-major, 4 chains per rank)
        assert liks == [-100.0, -99.0, -98.0, -97.0, -90.0, -10.0, -88.0, -87.0]
        assert id_sum == sum(range(128))       # the id created by rank 0 arrived intact

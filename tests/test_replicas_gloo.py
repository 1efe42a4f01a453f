"""Host-side replica logic over torch.distributed with the gloo backend, world_size 2, on CPU."""
import os
import socket

import numpy as np
import pytest

from instagraal_b200.replicas import ReplicaExchange, best_chain, exchange_pairs


def test_best_chain_and_exchange_are_deterministic():
    assert best_chain([-5.0, -3.0, -3.0, -9.0]) == 1
    sw = exchange_pairs([-10.0, -2.0, -8.0, -1.0], [1.0, 1.3, 1.6, 2.0], 0, np.zeros(4) + 1e-9)
    assert sw == [(0, 1), (2, 3)]
    assert exchange_pairs([-1.0, -2.0], [1.0, 2.0], 0, np.ones(2)) == []


def test_allgather_world_size_2_gloo():
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
    for rank, best, liks, ncs, first in sorted(res):
        assert best == 1 and liks == [-100.0, -90.0] and ncs == [7, 8] and first == [0, 1000]

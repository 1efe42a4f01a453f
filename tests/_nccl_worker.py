"""helper process for test_gpu_replicas.py::test_nccl_allgather_two_gpus: one rank (= one GPU) with 2 chains"""
import json
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, idfile = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    import numpy as np
    from instagraal_b200.cuda_lib_gl_single import sampler
    from instagraal_b200.replicas import ReplicaSet, nccl_unique_id
    from instagraal_b200.synth import WORKLOADS, make_level
    p8 = np.array([2.2354, 1.4933294, 0.06928191, -0.9384134, 2.0, 386.88467, 65.71848, 0.01698581], dtype=np.float32)
    level = make_level(WORKLOADS["toy"])
    first = sampler(*level.sampler_args(), device=rank)
    first.set_param_simu(p8)
    rs = ReplicaSet(first, 2, seeds=np.array([10 * rank + 1, 10 * rank + 2], dtype=np.uint64))
    if rank == 0:
        with open(idfile + ".tmp", "wb") as fh:
            fh.write(nccl_unique_id())
        os.replace(idfile + ".tmp", idfile)
    t0 = time.time()
    while not os.path.exists(idfile):
        assert time.time() - t0 < 120
        time.sleep(0.05)
    rs.init_comm(rank, world, open(idfile, "rb").read())
    rs.bomb(seed0=50 + 10 * rank)
    rng = np.random.RandomState(rank)
    frags = np.stack([rng.permutation(level.n_frags) for _ in range(2)]).astype(np.int32)
    out = rs.run_cycle(frags, 5, cycle=0)
    best, lik, nc = rs.allgather()
    st = rs.gathered_state(best)
    print("RESULT " + json.dumps(dict(rank=rank, best=best, lik=lik.tolist(), nc=nc.tolist(), own=[float(out[i][-1]["likelihood"]) for i in range(2)],
                                      best_state_crc=zlib.crc32(st.tobytes()))), flush=True)
    rs.free()


if __name__ == "__main__":
    main()

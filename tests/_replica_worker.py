"""helper process for test_replicas_gloo.py: one rank of a world_size-2 gloo group on CPU"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    import numpy as np
    import torch
    import torch.distributed as dist
    from instagraal_b200.replicas import best_chain
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the launcher's only job: hand the 128-byte NCCL id of rank 0 to every rank
    idt = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
    dist.broadcast(idt, 0)
    # what ig_allgather_best returns on every rank: the rank-major table of the chains' likelihoods
    n_local = 4
    mine = torch.tensor([-100.0 + 10 * rank + i for i in range(n_local)], dtype=torch.float64)
    if rank == 1:
        mine[1] = -10.0
    table = [torch.zeros(n_local, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(table, mine)
    lik = np.concatenate([t.numpy() for t in table])
    print("RESULT " + json.dumps([rank, best_chain(lik), lik.tolist(), int(idt.sum())]), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""helper process for test_replicas_gloo.py: one rank of a world_size-2 gloo group on CPU"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, port = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    import torch
    import torch.distributed as dist
    from instagraal_b200.replicas import ReplicaExchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class FakeSampler:
        likelihood_t = -100.0 + 10 * rank
        n_contigs = 7 + rank

    state = torch.arange(32, dtype=torch.int32) + 1000 * rank
    x = ReplicaExchange(FakeSampler(), dist, torch.device("cpu"), temperature=1.0 + rank, state_fn=lambda: state)
    best, liks, ncs = x.allgather()
    print("RESULT " + json.dumps([rank, best, liks.tolist(), ncs.tolist(), x.all_states[:, 0].tolist()]), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Replica chains across GPUs (a single chain is sequential, so the only natural
sharding is independent chains / tempering replicas, one or more per GPU).

One process per GPU (torchrun); each process owns one ``sampler`` (one chain).  Every
``gather_every`` steps the chains all-gather {likelihood, n_contigs, temperature, live scaffold}
over NCCL (NVLink/NVSwitch) -- ~64 B x NF per chain, latency-bound -- and every rank takes the same
deterministic decisions from the gathered table (best chain; optional replica-exchange swaps).
torch.distributed is plumbing only; the data path has no other collective.
"""
from __future__ import annotations

import numpy as np


class _DevBuf:
    """Expose a raw device pointer (the handle's live scaffold) through __cuda_array_interface__."""

    def __init__(self, ptr, n_int32):
        self.__cuda_array_interface__ = {"shape": (int(n_int32),), "typestr": "<i4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def best_chain(likelihoods):
    """Deterministic on every rank: highest likelihood, lowest rank on ties."""
    lik = np.asarray(likelihoods, dtype=np.float64)
    return int(np.flatnonzero(lik == lik.max())[0])


def exchange_pairs(likelihoods, temperatures, sweep, u):
    """Replica-exchange (parallel tempering) decisions for neighbouring temperature pairs
    (even pairs on even sweeps, odd pairs on odd sweeps), Metropolis on
    (1/T_i - 1/T_j) * (L_j - L_i) with the shared uniform draws ``u`` -- identical on every rank."""
    lik = np.asarray(likelihoods, dtype=np.float64)
    T = np.asarray(temperatures, dtype=np.float64)
    order = np.argsort(T, kind="stable")
    swaps = []
    for k in range(sweep % 2, len(order) - 1, 2):
        i, j = int(order[k]), int(order[k + 1])
        log_r = (1.0 / T[i] - 1.0 / T[j]) * (lik[j] - lik[i])
        if np.log(max(u[k], 1e-300)) < log_r:
            swaps.append((i, j))
    return swaps


class ReplicaExchange:
    def __init__(self, sampler, dist, device, temperature=1.0, state_fn=None):
        import torch
        self.torch = torch
        self.s = sampler
        self.dist = dist
        self.device = device
        self.world = dist.get_world_size()
        self.rank = dist.get_rank()
        self.temperature = float(temperature)
        self.state_fn = state_fn or self._device_state
        n = self.state_fn().numel()
        self.all_states = torch.empty((self.world, n), dtype=torch.int32, device=device)
        self.all_meta = torch.empty((self.world, 3), dtype=torch.float64, device=device)
        self.n_gathers = 0

    def _device_state(self):
        ptr, nbytes = self.s.device_state()
        return self.torch.as_tensor(_DevBuf(ptr, nbytes // 4), device=self.device)

    def allgather(self):
        """Returns (best rank, likelihoods[world], n_contigs[world]); all_states holds every chain's scaffold."""
        t = self.torch
        lik = float(self.s.likelihood_t) if self.s.likelihood_t is not None else float("-inf")
        meta = t.tensor([lik, float(self.s.n_contigs or 0), self.temperature], dtype=t.float64, device=self.device)
        self.dist.all_gather([self.all_meta[i] for i in range(self.world)], meta)
        self.dist.all_gather([self.all_states[i] for i in range(self.world)], self.state_fn().contiguous())
        m = self.all_meta.cpu().numpy()
        self.n_gathers += 1
        return best_chain(m[:, 0]), m[:, 0].copy(), m[:, 1].astype(np.int64)

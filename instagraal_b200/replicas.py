"""Replica chains (a single chain is sequential, so the only natural sharding of the 8 x B200 box is
independent chains with different seeds, one or more per GPU).

``ReplicaSet`` owns the chains of ONE process / GPU: the first one is a normal ``sampler`` (uploads the level), the others
are ``ig_clone``s that share its contacts in device memory.  ``run_cycle`` advances all of them together
(``ig_run_cycles_device_multi``: steps enqueued round-robin on the chains' streams, so a GPU that one yeast-scale chain
keeps < 25 % busy is filled by eight); ``allgather`` is the per-cycle exchange of {likelihood, n_contigs, live scaffold}
across the GPUs: ``ncclAllGather`` INSIDE the library, straight from device memory over NVLink (``ig_allgather_best``).
No PyTorch here: the launcher only has to hand every rank the 128-byte NCCL id created by rank 0
(``nccl_unique_id`` -> file / environment / any broadcast).
The live path of the reference has no temperature (fragment moves are an argmax, CL:1435-1446), so the chains differ by
their seeds only; every rank takes the same decision (best chain) from the gathered table.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L


def best_chain(likelihoods):
    """Deterministic on every rank: highest likelihood, lowest global chain index on ties (what ig_allgather_best returns)."""
    lik = np.asarray(likelihoods, dtype=np.float64)
    return int(np.flatnonzero(lik == lik.max())[0])


def nccl_unique_id():
    """128-byte id for ncclCommInitRank (call on rank 0, distribute to the other ranks)."""
    buf = (C.c_char * 128)()
    L.check(None, L.lib().ig_nccl_unique_id(buf), "ig_nccl_unique_id")
    return bytes(buf)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class ReplicaSet:
    def __init__(self, first_sampler, n_chains, seeds=None):
        """first_sampler: a constructed ``sampler`` (parameters already set); n_chains - 1 clones are made from it."""
        from .cuda_lib_gl_single import sampler
        self.chains = [first_sampler] + [sampler.clone_of(first_sampler) for _ in range(n_chains - 1)]
        self.n = n_chains
        self.seeds = np.ascontiguousarray(seeds if seeds is not None else np.arange(n_chains), dtype=np.uint64)
        self._handles = (C.c_void_p * n_chains)(*[c._h for c in self.chains])
        self.rank, self.n_ranks = 0, 1
        self._nccl = False
        self.nf = int(first_sampler.n_new_frags)
        self.gather_ms = 0.0
        self.n_gathers = 0
        # several chains on this GPU: each takes a share of the SMs so that their (latency-bound) kernels overlap
        self.gpu_share = min(n_chains, 4)
        for c in self.chains:
            c.set_gpu_share(self.gpu_share)

    def init_comm(self, rank=0, n_ranks=1, nccl_id=None):
        lead = self.chains[0]
        idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        L.check(lead._h, L.lib().ig_nccl_init(lead._h, int(rank), int(n_ranks), self.n, idbuf), "ig_nccl_init")
        self.rank, self.n_ranks, self._nccl = int(rank), int(n_ranks), True

    def bomb(self, seed0=0):
        for i, c in enumerate(self.chains):
            np.random.seed(int(seed0) + i)
            c.bomb_the_genome()

    def run_cycle(self, frags_per_chain, n_neighbours=5, cycle=0):
        """frags_per_chain: int32[n_chains, n_steps] visiting orders.  Returns the structured per-step records
        [n_chains, n_steps] (same fields as sampler.run_cycle_device)."""
        frags = np.ascontiguousarray(frags_per_chain, dtype=np.int32)
        assert frags.ndim == 2 and frags.shape[0] == self.n
        for c in self.chains:
            c._upload_neighbour_weights()
        n_steps = frags.shape[1]
        out = np.zeros((self.n, n_steps), dtype=L.CYCLE_DTYPE)
        rc = L.lib().ig_run_cycles_device_multi(self._handles, self.n, n_steps, _ptr(frags), int(n_neighbours), _ptr(self.seeds),
                                                int(cycle), _ptr(out))
        L.check(None, rc, "ig_run_cycles_device_multi")
        for i, c in enumerate(self.chains):
            c._after_cycle(out[i])
        return out

    def allgather(self):
        """-> (best global chain index, likelihood[n_ranks * n], n_contigs[n_ranks * n]); device ms accumulated in gather_ms"""
        if not self._nccl:
            self.init_comm()
        n_all = self.n_ranks * self.n
        lik = np.zeros(n_all, dtype=np.float64)
        nc = np.zeros(n_all, dtype=np.int32)
        best, ms = C.c_int32(0), C.c_float(0.0)
        lead = self.chains[0]
        L.check(lead._h, L.lib().ig_allgather_best(lead._h, self._handles, self.n, _ptr(lik), _ptr(nc), C.byref(best), C.byref(ms)),
                "ig_allgather_best")
        self.gather_ms += float(ms.value)
        self.n_gathers += 1
        return int(best.value), lik, nc

    def gathered_state(self, index):
        out = np.zeros((13, self.nf), dtype=np.int32)
        lead = self.chains[0]
        L.check(lead._h, L.lib().ig_get_gathered_state(lead._h, int(index), _ptr(out)), "ig_get_gathered_state")
        return out

    def free(self):
        for c in reversed(self.chains):
            c.free_gpu()

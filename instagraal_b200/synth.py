"""Synthetic in-silico assemblies + simulated Hi-C contacts, emitted as exactly the arrays the
reference ``sampler`` constructor takes (reference: simu_single.py:120-153 call site,
cuda_lib_gl_single.py:92-125 signature).

This replaces, for tests and benchmarks, the reference's on-disk pipeline
(pre.py -> pyramid_sparse.py -> simu_single.py), which needs h5py/Biopython/cooler and real
read pairs.  The layouts mirror:
  * ``level.load_data``            pyramid_sparse.py:1836-1849  (S_o_A_frags, 1-based id_c)
  * ``simulation.create_sub_frags`` simu_single.py:674-723      (float4 sub_frags_2_frags, int4 ids)
  * ``level.load_data`` trans mean pyramid_sparse.py:1875-1899  (mean_value_trans)

Contacts follow the reference's own polymer model (optim_rippe_curve_update.py:21-31,66-70):
lambda(s) = A * 0.53 * kuhn^-3 * (lm*s/kuhn)^slope for cis pairs on the TRUE chromosome layout,
plus a uniform trans/noise floor.  Everything is seeded.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

INT4 = np.dtype([("x", np.int32), ("y", np.int32), ("z", np.int32), ("w", np.int32)], align=True)
INT3 = np.dtype([("x", np.int32), ("y", np.int32), ("z", np.int32)], align=True)
INT2 = np.dtype([("x", np.int32), ("y", np.int32)], align=True)
FLOAT3 = np.dtype([("x", np.float32), ("y", np.float32), ("z", np.float32)], align=True)
FLOAT4 = np.dtype([("x", np.float32), ("y", np.float32), ("z", np.float32), ("w", np.float32)], align=True)

SOA_KEYS = ("pos", "sub_pos", "id_c", "start_bp", "len_bp", "sub_len", "circ", "id", "prev", "next",
            "l_cont", "sub_l_cont", "l_cont_bp", "n_accu", "rep", "activ", "id_d")


@dataclass
class SynthSpec:
    n_frags: int = 120            # level-L fragments
    n_contigs: int = 8            # initial contigs
    n_chrom: int = 2              # true chromosomes
    mean_sub_len_bp: float = 3300.0
    sigma_sub_len: float = 0.35   # log-normal sigma of sub-fragment length
    lambda1: float = 25.0         # expected contacts between adjacent sub-fragments
    slope: float = -1.5
    max_offset: int = 400         # band half-width (in sub-fragments) for cis sampling
    trans_per_row: float = 2.0    # mean number of uniform noise contacts per sub-fragment row
    diag_mean: float = 30.0       # self-contacts (diagonal; removed by the sampler)
    shuffle_contigs: bool = True  # initial contig order/orientation differs from the truth
    seed: int = 42


@dataclass
class LevelData:
    spec: SynthSpec
    n_frags: int
    n_sub_frags: int
    S_o_A_frags: dict
    S_o_A_sub_frags: dict
    np_sub_frags_2_frags: np.ndarray      # float4[NS]
    np_sub_frags_id: np.ndarray           # int4[NF]
    np_sub_frags_len_bp: np.ndarray       # float3[NF] (kb)
    np_sub_frags_accu: np.ndarray         # int3[NF]
    np_rep_sub_frags_id: np.ndarray       # int4[NF]
    sparse_matrix: sp.csr_matrix          # level L-1, upper incl. diagonal, int32
    sub_sampled_sparse_matrix: sp.csr_matrix  # level L, upper incl. diagonal, int32
    mean_value_trans: float
    true_order: np.ndarray = field(default=None)  # sub-frag ids along the true genome
    true_chrom: np.ndarray = field(default=None)  # chromosome of each entry of true_order

    def true_state(self):
        """int32[13, NF] scaffold of the TRUE assembly (one contig per chromosome, fragments in true
        order, reversed initial contigs carried as ori = -1): the fully assembled regime."""
        parent = self.np_sub_frags_2_frags["x"].astype(np.int64)
        fr = parent[self.true_order]
        keep = np.r_[True, fr[1:] != fr[:-1]]
        frags = fr[keep]                      # fragments in true order
        chrom = self.true_chrom[keep]
        sub_first = self.true_order[keep]     # first sub-fragment met for each fragment
        j_first = self.np_sub_frags_2_frags["w"].astype(np.int64)[sub_first]
        ori = np.where(j_first == 0, 1, -1)
        soa = self.S_o_A_frags
        sub_len = np.asarray(soa["sub_len"])[frags]
        ori = np.where(sub_len == 1, 1, ori)
        len_bp = np.asarray(soa["len_bp"])[frags]
        st = _soa(chrom + 1, len_bp, sub_len)
        out = np.zeros((13, self.n_frags), dtype=np.int32)
        names = ("pos", "sub_pos", "id_c", "start_bp", "len_bp", "sub_len", "circ", "prev", "next", "l_cont",
                 "sub_l_cont", "l_cont_bp")
        for i, k in enumerate(names):
            v = st[k]
            if k in ("prev", "next"):
                v = np.where(v >= 0, frags[np.clip(v, 0, None)], -1)
            out[i, frags] = v
        out[12, frags] = ori
        return out

    def sampler_args(self):
        """The 29 positional constructor arguments of the reference ``sampler``."""
        nf, ns = self.n_frags, self.n_sub_frags
        disp = np.zeros(nf, dtype=INT2)
        disp["x"] = np.arange(nf)
        disp["y"] = np.arange(nf) + 1
        sdisp = np.zeros(ns, dtype=INT2)
        sdisp["x"] = np.arange(ns)
        sdisp["y"] = np.arange(ns) + 1
        rng = np.random.RandomState(self.spec.seed + 12345)
        vel = rng.rand(nf, 4).astype(np.float32)
        pos = rng.rand(nf, 4).astype(np.float32)
        accu = np.asarray(self.np_sub_frags_accu["x"] + self.np_sub_frags_accu["y"] + self.np_sub_frags_accu["z"])
        return (
            True,                                   # use_rippe
            self.S_o_A_frags,
            np.arange(nf, dtype=np.int32),          # collector_id_repeats
            disp,                                   # frag_dispatcher
            [],                                     # id_frag_duplicated
            [],                                     # id_frags_blacklisted
            nf, nf, ns, ns,
            self.np_rep_sub_frags_id,
            self.sub_sampled_sparse_matrix,
            self.np_sub_frags_len_bp,
            self.np_sub_frags_id,
            self.np_sub_frags_accu,
            self.np_sub_frags_2_frags,
            np.float32(1.0),                        # mean_squared_frags_per_bin
            np.asmatrix(accu),                      # norm_vect_accu
            [], [],                                 # sub_candidates_dup, sub_candidates_output_data
            self.S_o_A_sub_frags,
            np.arange(ns, dtype=np.int32),          # sub_collector_id_repeats
            sdisp,                                  # sub_frag_dispatcher
            self.sparse_matrix,
            self.mean_value_trans,
            30, False, vel, pos,
        )


def _soa(contig_of, len_bp, sub_len):
    """Struct-of-arrays for consecutive fragments grouped by (1-based, non-decreasing) contig id."""
    n = len(contig_of)
    contig_of = np.asarray(contig_of, dtype=np.int64)
    len_bp = np.asarray(len_bp, dtype=np.int64)
    sub_len = np.asarray(sub_len, dtype=np.int64)
    first = np.r_[True, contig_of[1:] != contig_of[:-1]]
    start_idx = np.flatnonzero(first)
    cont_idx = np.cumsum(first) - 1
    idx = np.arange(n)
    pos = idx - start_idx[cont_idx]
    cs_bp = np.cumsum(len_bp) - len_bp
    cs_sub = np.cumsum(sub_len) - sub_len
    start_bp = cs_bp - cs_bp[start_idx][cont_idx]
    sub_pos = cs_sub - cs_sub[start_idx][cont_idx]
    l_cont = np.diff(np.r_[start_idx, n])[cont_idx]
    l_cont_bp = np.add.reduceat(len_bp, start_idx)[cont_idx]
    sub_l_cont = np.add.reduceat(sub_len, start_idx)[cont_idx]
    last = np.r_[first[1:], True]
    prev = np.where(first, -1, idx - 1)
    nxt = np.where(last, -1, idx + 1)
    d = {
        "pos": pos, "sub_pos": sub_pos, "id_c": contig_of, "start_bp": start_bp, "len_bp": len_bp,
        "sub_len": sub_len, "circ": np.zeros(n), "id": idx, "prev": prev, "next": nxt,
        "l_cont": l_cont, "sub_l_cont": sub_l_cont, "l_cont_bp": l_cont_bp,
        "n_accu": sub_len, "rep": np.zeros(n), "activ": np.ones(n), "id_d": idx,
    }
    return {k: np.ascontiguousarray(v, dtype=np.int32) for k, v in d.items()}


def load_geometry(path):
    """Fragment geometry of a real digested assembly (a fixture generated from the reference's tests/data contigs
    by the test tooling, see tests/golden/): fragments per contig, sub-fragments per fragment, sub-fragment lengths, true layout."""
    z = np.load(path)
    order = np.lexsort((z["contig_start"], z["contig_chrom"]))   # contigs along the true genome
    return {"sizes": z["frags_per_contig"].astype(np.int64), "sub_len": z["frag_nsub"].astype(np.int64),
            "sub_len_bp": z["sub_len_bp"].astype(np.int64), "order": order.astype(np.int64),
            "chrom_of_rank": z["contig_chrom"][order].astype(np.int64)}


def make_level(spec: SynthSpec, geometry=None) -> LevelData:
    rng = np.random.RandomState(spec.seed)
    if geometry is not None:
        sizes = np.asarray(geometry["sizes"], dtype=np.int64)
        nf, nc = int(sizes.sum()), int(sizes.size)
        contig_of = np.repeat(np.arange(1, nc + 1), sizes)
        sub_len = np.asarray(geometry["sub_len"], dtype=np.int64)
        ns = int(sub_len.sum())
        sub_len_bp = np.asarray(geometry["sub_len_bp"], dtype=np.int64)
        assert sub_len.size == nf and sub_len_bp.size == ns
    else:
        nf, nc = int(spec.n_frags), int(spec.n_contigs)
        assert nc >= 1 and nf >= nc
        # ---- initial contigs: log-normal lengths (in fragments), each >= 1
        w = rng.lognormal(0.0, 0.6, nc)
        sizes = np.maximum(1, np.floor(w / w.sum() * nf)).astype(np.int64)
        while sizes.sum() > nf:
            sizes[np.argmax(sizes)] -= 1
        sizes[np.argmax(sizes)] += nf - sizes.sum()
        contig_of = np.repeat(np.arange(1, nc + 1), sizes)
        # ---- sub-fragments: 3 per fragment, last fragment of each contig 1..3
        sub_len = np.full(nf, 3, dtype=np.int64)
        last_idx = np.cumsum(sizes) - 1
        sub_len[last_idx] = rng.randint(1, 4, nc)
        ns = int(sub_len.sum())
        sub_len_bp = np.maximum(200, rng.lognormal(np.log(spec.mean_sub_len_bp), spec.sigma_sub_len, ns)).astype(np.int64)
    parent = np.repeat(np.arange(nf), sub_len)
    first_sub = np.cumsum(sub_len) - sub_len
    j_in_parent = np.arange(ns) - first_sub[parent]
    len_bp = np.add.reduceat(sub_len_bp, first_sub)
    soa = _soa(contig_of, len_bp, sub_len)
    sub_soa = _soa(contig_of[parent], sub_len_bp, np.ones(ns, dtype=np.int64))

    # ---- float4 sub_frags_2_frags (simu_single.py:703-717), float32 arithmetic like the reference
    kb = (sub_len_bp.astype(np.float32) / np.float32(1000.0)).astype(np.float32)
    s2f = np.zeros(ns, dtype=FLOAT4)
    ids4 = np.zeros(nf, dtype=INT4)
    len3 = np.zeros(nf, dtype=FLOAT3)
    accu3 = np.zeros(nf, dtype=INT3)
    wat = np.zeros(ns, dtype=np.float32)
    cri = np.zeros(ns, dtype=np.float32)
    for n_sub in (1, 2, 3):
        fr = np.flatnonzero(sub_len == n_sub)
        if fr.size == 0:
            continue
        base = first_sub[fr]
        L = np.stack([kb[base + j] for j in range(n_sub)], axis=1)  # (m, n_sub) float32
        for j in range(n_sub):
            acc_w = np.zeros(fr.size, dtype=np.float32)
            for t in range(0, j):
                acc_w = (acc_w + L[:, t]).astype(np.float32)
            acc_c = np.zeros(fr.size, dtype=np.float32)
            for t in range(n_sub - 1, j, -1):
                acc_c = (acc_c + L[:, t]).astype(np.float32)
            half = (L[:, j] / np.float32(2.0)).astype(np.float32)
            wat[base + j] = (acc_w + half).astype(np.float32)
            cri[base + j] = (acc_c + half).astype(np.float32)
    s2f["x"] = parent.astype(np.float32)
    s2f["y"] = wat
    s2f["z"] = cri
    s2f["w"] = j_in_parent.astype(np.float32)
    for j, key in enumerate(("x", "y", "z")):
        has = sub_len > j
        ids4[key][has] = (first_sub[has] + j).astype(np.int32)
        len3[key][has] = kb[first_sub[has] + j]
        accu3[key][has] = 1
    ids4["w"] = sub_len.astype(np.int32)
    rep4 = ids4.copy()

    # ---- TRUE genome: contigs are pieces of n_chrom chromosomes; the initial order/orientation
    #      (what the sampler sees) is a shuffled / partly reversed version of the truth.
    order = np.arange(nc)
    flip = np.zeros(nc, dtype=bool)
    if geometry is not None:
        order = np.asarray(geometry["order"], dtype=np.int64)
        chrom_of_contig = np.asarray(geometry["chrom_of_rank"], dtype=np.int64)
    else:
        if spec.shuffle_contigs:
            order = rng.permutation(nc)
            flip = rng.rand(nc) < 0.5
        chrom_of_contig = np.sort(rng.randint(0, spec.n_chrom, nc))  # in true order
    sub_first_of_contig = np.cumsum(np.add.reduceat(sub_len, np.cumsum(sizes) - sizes)) - np.add.reduceat(
        sub_len, np.cumsum(sizes) - sizes)
    sub_count_of_contig = np.add.reduceat(sub_len, np.cumsum(sizes) - sizes)
    true_ids = []
    true_chrom = []
    for rank, c in enumerate(order):
        ids = np.arange(sub_first_of_contig[c], sub_first_of_contig[c] + sub_count_of_contig[c])
        if flip[c]:
            ids = ids[::-1]
        true_ids.append(ids)
        true_chrom.append(np.full(ids.size, chrom_of_contig[rank]))
    true_ids = np.concatenate(true_ids)
    true_chrom = np.concatenate(true_chrom)
    tl = sub_len_bp[true_ids].astype(np.float64) / 1000.0
    centre = np.cumsum(tl) - tl / 2.0  # kb along the concatenated true genome

    rows, cols, vals = [], [], []
    # cis band: Poisson(lambda1 * (s/s1)^slope) with s1 = mean adjacent distance
    s1 = float(np.mean(np.abs(np.diff(centre)))) if ns > 1 else 1.0
    max_off = min(int(spec.max_offset), ns - 1)
    for k in range(1, max_off + 1):
        same = true_chrom[k:] == true_chrom[:-k]
        if not same.any():
            continue
        s = (centre[k:] - centre[:-k])[same]
        lam = spec.lambda1 * np.power(s / s1, spec.slope)
        cnt = rng.poisson(lam)
        nz = cnt > 0
        if nz.any():
            a = true_ids[:-k][same][nz]
            b = true_ids[k:][same][nz]
            rows.append(np.minimum(a, b))
            cols.append(np.maximum(a, b))
            vals.append(cnt[nz])
    # uniform noise / trans floor
    n_noise = rng.poisson(spec.trans_per_row * ns)
    if n_noise > 0 and ns > 1:
        a = rng.randint(0, ns, n_noise)
        b = rng.randint(0, ns, n_noise)
        keep = a != b
        rows.append(np.minimum(a, b)[keep])
        cols.append(np.maximum(a, b)[keep])
        vals.append(np.ones(int(keep.sum()), dtype=np.int64))
    # diagonal
    dg = rng.poisson(spec.diag_mean, ns)
    rows.append(np.arange(ns)[dg > 0])
    cols.append(np.arange(ns)[dg > 0])
    vals.append(dg[dg > 0])
    rows = np.concatenate(rows).astype(np.int64)
    cols = np.concatenate(cols).astype(np.int64)
    vals = np.concatenate(vals).astype(np.int64)
    key = rows * ns + cols
    o = np.argsort(key, kind="stable")
    key, vals = key[o], vals[o]
    uniq = np.r_[True, key[1:] != key[:-1]]
    st = np.flatnonzero(uniq)
    vals = np.add.reduceat(vals, st)
    key = key[st]
    rows = (key // ns).astype(np.int32)
    cols = (key % ns).astype(np.int32)
    vals = vals.astype(np.int32)
    mat = sp.csr_matrix((vals, (rows, cols)), shape=(ns, ns), dtype=np.int32)
    mat.sum_duplicates()
    mat.sort_indices()
    # level L = x3 binning of level L-1 (sum over parent fragments), upper incl. diagonal
    pr, pc = parent[rows], parent[cols]
    lo, hi = np.minimum(pr, pc), np.maximum(pr, pc)
    sub_mat = sp.coo_matrix((vals.astype(np.int64), (lo, hi)), shape=(nf, nf)).tocsr()
    sub_mat.sum_duplicates()
    sub_mat.sort_indices()
    sub_mat = sub_mat.astype(np.int32)

    # mean trans value at level L-1 w.r.t. the INITIAL contigs (pyramid_sparse.py:1875-1899)
    cr, cc = contig_of[parent][rows], contig_of[parent][cols]
    sub_sizes = sub_count_of_contig.astype(np.float64)
    # the reference sums full rows of each contig minus its intra block on the upper-stored matrix
    total_trans = float(vals[cr != cc].sum())
    n_tot_intra = float(np.sum(sub_sizes * (sub_sizes - 1) / 2))
    n_tot = ns * (ns - 1) / 2 - n_tot_intra
    mvt = total_trans / np.float32(n_tot) if n_tot > 0 else 0.0
    if not np.isfinite(mvt) or mvt <= 0:
        mvt = float(vals.min()) / 10.0
    return LevelData(
        spec=spec, n_frags=nf, n_sub_frags=ns, S_o_A_frags=soa, S_o_A_sub_frags=sub_soa,
        np_sub_frags_2_frags=s2f, np_sub_frags_id=ids4, np_sub_frags_len_bp=len3, np_sub_frags_accu=accu3,
        np_rep_sub_frags_id=rep4, sparse_matrix=mat, sub_sampled_sparse_matrix=sub_mat,
        mean_value_trans=float(mvt), true_order=true_ids, true_chrom=true_chrom,
    )


def workload_params(level):
    """float32[8] param_simu (kuhn, lm, c1, slope, d, d_max, fact, v_inter; KA:91-100) the contacts of `level` were
    simulated with (optim_rippe_curve_update.py:66-70 defaults; fact / v_inter from the workload's spec)."""
    spec = level.spec
    kuhn, lm, slope = 50.0, 9.6, -1.5
    c1 = np.float32(0.53 * (lm / kuhn) ** slope * kuhn ** -3)
    s1 = float(level.S_o_A_sub_frags["len_bp"].mean()) / 1000.0
    fact = spec.lambda1 / (float(c1) * s1 ** slope)
    ns = level.n_sub_frags
    v_inter = max(spec.trans_per_row * 2.0 / ns, 1e-6) / 10.0
    d_max = (v_inter / (float(c1) * fact)) ** (1.0 / slope)
    return np.array([kuhn, lm, c1, slope, 2.0, d_max, fact, v_inter], dtype=np.float32)


# Named workloads.  T = toy/yeast-like level 4; Y3 = yeast level 3;
# G = ~1 Gb synthetic (1e5 fragments, ~3e5 sub-fragments, ~1e8 contacts).
YEAST_TOY_GEOMETRY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "yeast_toy_geometry.npz")


def make_workload(name):
    """Named workload -> LevelData.  "yeast_toy" = BASELINE.json configs[0]: the reference's tests/data contigs
    digested with DpnII + HinfI and binned to level 4 (geometry fixture), contacts simulated with seed 0."""
    if name == "yeast_toy":
        return make_level(WORKLOADS["yeast_toy"], load_geometry(YEAST_TOY_GEOMETRY))
    return make_level(WORKLOADS[name])


WORKLOADS = {
    "yeast_toy": SynthSpec(max_offset=800, lambda1=300.0, trans_per_row=40.0, seed=0),
    "micro": SynthSpec(n_frags=40, n_contigs=5, n_chrom=2, max_offset=60, lambda1=20.0, seed=1),
    "toy": SynthSpec(n_frags=150, n_contigs=10, n_chrom=3, max_offset=200, lambda1=25.0, seed=42),
    "T": SynthSpec(n_frags=900, n_contigs=146, n_chrom=16, max_offset=800, lambda1=300.0, trans_per_row=40.0, seed=42),
    "Y3": SynthSpec(n_frags=2700, n_contigs=146, n_chrom=16, max_offset=1500, lambda1=100.0, trans_per_row=40.0, seed=42),
    "G": SynthSpec(n_frags=100000, n_contigs=2000, n_chrom=20, max_offset=2500, lambda1=420.0,
                   trans_per_row=150.0, seed=7),
}

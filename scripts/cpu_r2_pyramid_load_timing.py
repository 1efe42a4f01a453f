"""CPU timing of the pyramid load side against the live reference classes (build container only: needs /root/reference).
   python scripts/cpu_r2_pyramid_load_timing.py  ->  profiles/r2_pyramid_load_timing.json (numbers copied by hand)"""
import os, sys, time, types
import numpy as np
sys.path.insert(0, "/root/repo")
OUT = "/tmp/pl/pyr"
rng = np.random.RandomState(3)
n_contigs, n_levels = 2000, 2
sizes = np.maximum(1, rng.lognormal(3.6, 0.8, n_contigs).astype(int))      # ~ 1e5 fragments
total = int(sizes.sum())
def write_level(lvl, sizes, sub_sizes=None):
    d = os.path.join(OUT, "level_%d" % lvl); os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "%d_fragments_list.txt" % lvl), "w") as f:
        f.write("id\tchrom\tstart_pos\tend_pos\tsize\tgc_content\taccu_frag\tinit_frag_start\tinit_frag_end" + ("\tsub_frag_start\tsub_frag_end" if lvl else "") + "\n")
        sub0 = 1
        for c, n in enumerate(sizes):
            ends = np.cumsum(rng.randint(200, 4000, n)); starts = ends - np.diff(np.r_[0, ends])
            for i in range(n):
                row = [i + 1, "ctg%05d" % c, starts[i], ends[i], ends[i] - starts[i], 0.5, 3 if lvl else 1, 1, 1]
                if lvl:
                    k = 3 if i < n - 1 else max(1, int(sub_sizes[c]) - 3 * (n - 1)); row += [sub0, sub0 + k - 1]; sub0 += k
                f.write("\t".join(str(x) for x in row) + "\n")
            if lvl: sub0 = sub0
    with open(os.path.join(d, "%d_contig_info.txt" % lvl), "w") as f:
        f.write("contig\tlength_kb\tn_frags\tcumul_length\n")
    return int(np.sum(sizes))
lvl1_sizes = (sizes + 2) // 3
n0 = write_level(0, sizes); n1 = write_level(1, lvl1_sizes, sizes)
with open(os.path.join(OUT, "level_0", "0_sub_2_super_index_frag.txt"), "w") as f:
    f.write("current_id\tsuper_id\n")
    off0 = np.r_[0, np.cumsum(sizes)]; off1 = np.r_[0, np.cumsum(lvl1_sizes)]
    for c in range(n_contigs):
        for i in range(sizes[c]):
            f.write("%d\t%d\n" % (off0[c] + i + 1, off1[c] + i // 3 + 1))
def contacts(n, m):
    a = rng.randint(0, n, m); b = np.where(rng.rand(m) < 0.8, np.clip(a + rng.randint(0, 12, m), 0, n - 1), rng.randint(0, n, m))
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = np.unique(lo.astype(np.int64) * n + hi)
    return np.stack([key // n, key % n, rng.randint(1, 20, key.size)]).astype(np.int32)
data = {"0": {"data": contacts(n0, 3000000), "nfrags": np.array([[n0]], np.int32)}, "1": {"data": contacts(n1, 1500000), "nfrags": np.array([[n1]], np.int32)}}
print("fragments", n0, n1, "contacts", data["0"]["data"].shape[1], data["1"]["data"].shape[1], flush=True)
class F(dict):
    def close(self): pass
store = F(data)
# ours
from instagraal_b200.pyramid_load import pyramid
t0 = time.time(); p = pyramid(OUT, n_levels, data=store); t1 = time.time(); l1 = p.get_level(1); t2 = time.time(); l0 = p.get_level(0); t3 = time.time()
print("ours: pyramid %.2f s, level 1 load_data %.2f s, level 0 load_data %.2f s" % (t1 - t0, t2 - t1, t3 - t2), flush=True)
# reference
h5 = types.ModuleType("h5py"); h5.File = lambda path, mode="a": store; sys.modules["h5py"] = h5
sys.path.insert(0, "/root/repo/oracle/ref_harness"); sys.path.insert(0, "/root/reference/src")
os.chdir("/tmp/pl")
import logging; logging.disable(logging.CRITICAL)
import instagraal.pyramid_sparse as PS
t0 = time.time(); pr = PS.pyramid(OUT, n_levels); t1 = time.time(); r1 = pr.get_level(1); t2 = time.time(); r0 = pr.get_level(0); t3 = time.time()
print("reference: pyramid %.2f s, level 1 load_data %.2f s, level 0 load_data %.2f s" % (t1 - t0, t2 - t1, t3 - t2), flush=True)
for a, b in ((l1, r1), (l0, r0)):
    assert all(np.array_equal(a.S_o_A_frags[k], b.S_o_A_frags[k]) for k in b.S_o_A_frags) and float(a.mean_value_trans) == float(b.mean_value_trans)
print("outputs identical")

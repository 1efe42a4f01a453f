"""ORACLE (test infrastructure only -- never imported by the product path).

NumPy restatement of the reference's scaffold-mutation kernels, one function per kernel,
operating on whole struct-of-arrays copies exactly like the reference does (one thread per
fragment -> one vectorised mask per branch).  All arithmetic is int32/int64-exact.

Every function cites the kernel it follows in /root/reference/src/instagraal/kernels/
kernel_sparse_adapt.cu (abbreviated KA) and the Python launch site in
cuda_lib_gl_single.py (CL).  Pinned against the reference itself (run through the CPU
emulation in oracle/ref_harness) by tests/test_oracle_golden.py.

The repeat machinery is inert in the reference (rep=0, activ=1, id_d=id), so the
``activ`` guards of the kernels are always true and are not restated.
"""
from __future__ import annotations

import numpy as np

FIELDS = ("pos", "sub_pos", "id_c", "start_bp", "len_bp", "sub_len", "circ", "prev", "next",
          "l_cont", "sub_l_cont", "l_cont_bp", "ori")
ALL17 = ("pos", "sub_pos", "id_c", "start_bp", "len_bp", "sub_len", "circ", "id", "prev", "next",
         "l_cont", "sub_l_cont", "l_cont_bp", "ori", "rep", "activ", "id_d")


def copy_state(s):
    return {k: v.copy() for k, v in s.items()}


def state_from_soa(soa):
    """Initial live scaffold (CL:521-549): data arrays + ori = +1."""
    n = len(soa["pos"])
    s = {k: np.array(soa[k], dtype=np.int32) for k in FIELDS if k != "ori"}
    s["ori"] = np.ones(n, dtype=np.int32)
    return s


def _piv(s, f):
    return {k: int(s[k][f]) for k in FIELDS}


def _set(o, mask, **kw):
    for k, v in kw.items():
        if np.isscalar(v):
            o[k][mask] = v
        else:
            o[k][mask] = np.asarray(v)[mask]


# --------------------------------------------------------------------------------------------
def flip_frag(s, a):
    """KA:612-670."""
    o = copy_state(s)
    o["ori"][a] = -o["ori"][a]
    return o


def pop_out_frag(s, a, max_id):
    """KA:737-1078: eject fragment ``a`` into the singleton contig ``max_id+1``."""
    o = copy_state(s)
    A = _piv(s, a)
    n = len(s["pos"])
    idx = np.arange(n)
    l = A["l_cont"]
    if l < 2:
        return o
    inc = s["id_c"] == A["id_c"]
    lt = inc & (s["pos"] < A["pos"])
    eq = inc & (s["pos"] == A["pos"])
    gt = inc & (s["pos"] > A["pos"])
    shrink = dict(l_cont=s["l_cont"] - 1, sub_l_cont=s["sub_l_cont"] - A["sub_len"],
                  l_cont_bp=s["l_cont_bp"] - A["len_bp"])
    single = dict(pos=0, sub_pos=0, id_c=max_id + 1, start_bp=0, circ=0, ori=1, prev=-1, next=-1,
                  l_cont=1, sub_l_cont=s["sub_len"], l_cont_bp=s["len_bp"])
    if l > 2:
        prev_lt = np.where((idx == A["next"]) & (A["circ"] == 1), A["prev"], s["prev"])
        next_lt = np.where(s["pos"] == A["pos"] - 1, A["next"], s["next"])
        _set(o, lt, prev=prev_lt, next=next_lt, **shrink)
        _set(o, eq, **single)
        prev_gt = np.where(s["pos"] == A["pos"] + 1, A["prev"], s["prev"])
        next_gt = np.where((idx == A["prev"]) & (A["circ"] == 1), A["next"], s["next"])
        _set(o, gt, pos=s["pos"] - 1, sub_pos=s["sub_pos"] - A["sub_len"],
             start_bp=s["start_bp"] - A["len_bp"], prev=prev_gt, next=next_gt, **shrink)
    else:  # l == 2
        _set(o, lt, circ=0, prev=-1, next=-1, **shrink)
        _set(o, eq, **single)
        _set(o, gt, pos=s["pos"] - 1, sub_pos=s["sub_pos"] - A["sub_len"],
             start_bp=s["start_bp"] - A["len_bp"], circ=0, prev=-1, next=-1, **shrink)
    return o


def pop_in_frag_1(s, a, b, max_id, ori_new):
    """KA:1081-1371: re-insert ``a`` immediately LEFT of ``b`` and cut the contig before it."""
    o = copy_state(s)
    A, B = _piv(s, a), _piv(s, b)
    n = len(s["pos"])
    idx = np.arange(n)
    inc = (s["id_c"] == B["id_c"]) & (idx != a)
    lt = inc & (s["pos"] < B["pos"])
    eq = inc & (s["pos"] == B["pos"])
    gt = inc & (s["pos"] > B["pos"])
    if B["circ"] == 0:
        tail = dict(l_cont=B["l_cont"] - B["pos"] + 1,
                    sub_l_cont=B["sub_l_cont"] - B["sub_pos"] + A["sub_len"],
                    l_cont_bp=B["l_cont_bp"] - B["start_bp"] + A["len_bp"])
        _set(o, lt, id_c=B["id_c"], circ=0, next=np.where(s["pos"] == B["pos"] - 1, -1, s["next"]),
             l_cont=B["pos"], sub_l_cont=B["sub_pos"], l_cont_bp=B["start_bp"])
        _set(o, eq, pos=1, sub_pos=A["sub_len"], id_c=max_id + 1, start_bp=A["len_bp"], circ=0,
             ori=B["ori"], prev=a, next=B["next"], **tail)
        _set(o, gt, pos=s["pos"] - B["pos"] + 1, sub_pos=s["sub_pos"] - B["sub_pos"] + A["sub_len"],
             id_c=max_id + 1, start_bp=s["start_bp"] - B["start_bp"] + A["len_bp"], circ=0, **tail)
        amask = idx == a
        _set(o, amask, pos=0, sub_pos=0, start_bp=0, circ=0, ori=ori_new, prev=-1, next=b,
             id_c=max_id + 1, **tail)
    else:
        grow = dict(l_cont=B["l_cont"] + 1, sub_l_cont=B["sub_l_cont"] + A["sub_len"],
                    l_cont_bp=B["l_cont_bp"] + A["len_bp"])
        _set(o, lt, pos=B["l_cont"] - B["pos"] + s["pos"] + 1,
             sub_pos=B["sub_l_cont"] - B["sub_pos"] + s["sub_pos"] + A["sub_len"], id_c=B["id_c"],
             start_bp=B["l_cont_bp"] - B["start_bp"] + s["start_bp"] + A["len_bp"], circ=0,
             next=np.where(s["pos"] == B["pos"] - 1, -1, s["next"]), **grow)
        _set(o, eq, pos=1, sub_pos=A["sub_len"], id_c=B["id_c"], start_bp=A["len_bp"],
             len_bp=B["len_bp"], sub_len=B["sub_len"], circ=0, ori=B["ori"], prev=a, next=B["next"], **grow)
        _set(o, gt, pos=s["pos"] - B["pos"] + 1, sub_pos=s["sub_pos"] - B["sub_pos"] + A["sub_len"],
             id_c=B["id_c"], start_bp=s["start_bp"] - B["start_bp"] + A["len_bp"], circ=0,
             next=np.where(idx == B["prev"], -1, s["next"]), **grow)
        amask = idx == a
        _set(o, amask, pos=0, sub_pos=0, start_bp=0, circ=0, ori=ori_new, prev=-1, next=b,
             id_c=B["id_c"], **grow)
    return o


def pop_in_frag_2(s, a, b, max_id, ori_new):
    """KA:1373-1686: re-insert ``a`` immediately RIGHT of ``b`` and cut the contig after it."""
    o = copy_state(s)
    A, B = _piv(s, a), _piv(s, b)
    n = len(s["pos"])
    idx = np.arange(n)
    inc = (s["id_c"] == B["id_c"]) & (idx != a)
    lt = inc & (s["pos"] < B["pos"])
    eq = inc & (s["pos"] == B["pos"])
    gt = inc & (s["pos"] > B["pos"])
    amask = idx == a
    if B["circ"] == 0:
        head = dict(l_cont=B["pos"] + 2, l_cont_bp=B["start_bp"] + B["len_bp"] + A["len_bp"],
                    sub_l_cont=B["sub_pos"] + B["sub_len"] + A["sub_len"])
        _set(o, lt, id_c=B["id_c"], circ=0, **head)
        _set(o, eq, id_c=B["id_c"], circ=0, ori=B["ori"], prev=B["prev"], next=a, **head)
        _set(o, gt, pos=s["pos"] - (B["pos"] + 1), sub_pos=s["sub_pos"] - (B["sub_pos"] + B["sub_len"]),
             id_c=max_id + 1, start_bp=s["start_bp"] - (B["start_bp"] + B["len_bp"]), circ=0,
             prev=np.where(s["pos"] == B["pos"] + 1, -1, s["prev"]),
             l_cont=B["l_cont"] - (B["pos"] + 1), l_cont_bp=B["l_cont_bp"] - (B["start_bp"] + B["len_bp"]),
             sub_l_cont=B["sub_l_cont"] - (B["sub_pos"] + B["sub_len"]))
        _set(o, amask, pos=B["pos"] + 1, sub_pos=B["sub_pos"] + B["sub_len"], id_c=B["id_c"],
             start_bp=B["start_bp"] + B["len_bp"], circ=0, ori=ori_new, prev=b, next=-1, **head)
    else:
        grow = dict(l_cont=B["l_cont"] + 1, sub_l_cont=B["sub_l_cont"] + A["sub_len"],
                    l_cont_bp=B["l_cont_bp"] + A["len_bp"])
        rot = B["l_cont"] - (B["pos"] + 1)
        srot = B["sub_l_cont"] - (B["sub_pos"] + B["sub_len"])
        brot = B["l_cont_bp"] - (B["start_bp"] + B["len_bp"])
        _set(o, lt, pos=rot + s["pos"], sub_pos=srot + s["sub_pos"], id_c=B["id_c"],
             start_bp=brot + s["start_bp"], circ=0,
             prev=np.where(idx == B["next"], -1, s["prev"]), **grow)
        _set(o, eq, pos=rot + B["pos"], sub_pos=srot + B["sub_pos"], id_c=B["id_c"],
             start_bp=brot + B["start_bp"], len_bp=B["len_bp"], sub_len=B["sub_len"], circ=0,
             prev=B["prev"], next=a, **grow)
        _set(o, gt, pos=s["pos"] - (B["pos"] + 1), sub_pos=s["sub_pos"] - (B["sub_pos"] + B["sub_len"]),
             id_c=B["id_c"], start_bp=s["start_bp"] - (B["start_bp"] + B["len_bp"]), circ=0,
             prev=np.where(s["pos"] == B["pos"] + 1, -1, s["prev"]), **grow)
        _set(o, amask, pos=rot + B["pos"] + 1, sub_pos=srot + B["sub_pos"] + B["sub_len"], id_c=B["id_c"],
             start_bp=brot + B["start_bp"] + B["len_bp"], circ=0, ori=ori_new, prev=b, next=-1, **grow)
    return o


def pop_in_frag_3(s, a, b, max_id, ori_new):
    """KA:1688-1905: re-insert ``a`` immediately RIGHT of ``b`` (no cut, circularity kept)."""
    o = copy_state(s)
    A, B = _piv(s, a), _piv(s, b)
    n = len(s["pos"])
    idx = np.arange(n)
    inc = (s["id_c"] == B["id_c"]) & (idx != a)
    lt = inc & (s["pos"] < B["pos"])
    eq = inc & (s["pos"] == B["pos"])
    gt = inc & (s["pos"] > B["pos"])
    grow = dict(l_cont=B["l_cont"] + 1, sub_l_cont=B["sub_l_cont"] + A["sub_len"],
                l_cont_bp=B["l_cont_bp"] + A["len_bp"])
    _set(o, lt, id_c=B["id_c"], circ=B["circ"],
         prev=np.where((idx == B["next"]) & (B["circ"] == 1), a, s["prev"]), **grow)
    _set(o, eq, id_c=B["id_c"], circ=B["circ"], ori=B["ori"], next=a, **grow)
    _set(o, gt, pos=s["pos"] + 1, sub_pos=s["sub_pos"] + A["sub_len"], id_c=B["id_c"],
         start_bp=s["start_bp"] + A["len_bp"], circ=B["circ"],
         prev=np.where(s["pos"] == B["pos"] + 1, a, s["prev"]), **grow)
    _set(o, idx == a, pos=B["pos"] + 1, sub_pos=B["sub_pos"] + B["sub_len"], id_c=B["id_c"],
         start_bp=B["start_bp"] + B["len_bp"], circ=B["circ"], ori=ori_new, prev=b, next=B["next"], **grow)
    return o


def split_contig(s, f, upstream, max_id):
    """KA:2979-3365: cut the contig of ``f`` before (upstream=1) or after (upstream=0) it; circular
    contigs are rotated open instead (no new id)."""
    o = copy_state(s)
    F = _piv(s, f)
    if F["l_cont"] <= 1:
        return o
    n = len(s["pos"])
    idx = np.arange(n)
    inc = s["id_c"] == F["id_c"]
    lt = inc & (s["pos"] < F["pos"])
    eq = inc & (s["pos"] == F["pos"])
    gt = inc & (s["pos"] > F["pos"])
    if F["circ"] == 0:
        if upstream == 1:
            tail = dict(l_cont=F["l_cont"] - F["pos"], l_cont_bp=F["l_cont_bp"] - F["start_bp"],
                        sub_l_cont=F["sub_l_cont"] - F["sub_pos"])
            _set(o, lt, circ=0, next=np.where(s["pos"] == F["pos"] - 1, -1, s["next"]),
                 l_cont=F["pos"], l_cont_bp=F["start_bp"], sub_l_cont=F["sub_pos"])
            _set(o, eq, pos=0, sub_pos=0, id_c=max_id + 1, start_bp=0, circ=0, prev=-1, next=F["next"], **tail)
            _set(o, gt, pos=s["pos"] - F["pos"], sub_pos=s["sub_pos"] - F["sub_pos"], id_c=max_id + 1,
                 start_bp=s["start_bp"] - F["start_bp"], circ=0, **tail)
        else:
            head = dict(l_cont=F["pos"] + 1, l_cont_bp=F["start_bp"] + F["len_bp"],
                        sub_l_cont=F["sub_pos"] + F["sub_len"])
            _set(o, lt, circ=0, **head)
            _set(o, eq, circ=0, prev=F["prev"], next=-1, **head)
            _set(o, gt, pos=s["pos"] - (F["pos"] + 1), sub_pos=s["sub_pos"] - (F["sub_pos"] + F["sub_len"]),
                 id_c=max_id + 1, start_bp=s["start_bp"] - (F["start_bp"] + F["len_bp"]), circ=0,
                 prev=np.where(s["pos"] == F["pos"] + 1, -1, s["prev"]),
                 l_cont=F["l_cont"] - (F["pos"] + 1), l_cont_bp=F["l_cont_bp"] - (F["start_bp"] + F["len_bp"]),
                 sub_l_cont=F["sub_l_cont"] - (F["sub_pos"] + F["sub_len"]))
    else:
        keep = dict(l_cont=F["l_cont"], l_cont_bp=F["l_cont_bp"], sub_l_cont=F["sub_l_cont"])
        if upstream == 1:
            _set(o, lt, pos=F["l_cont"] - F["pos"] + s["pos"], sub_pos=F["sub_l_cont"] - F["sub_pos"] + s["sub_pos"],
                 start_bp=F["l_cont_bp"] - F["start_bp"] + s["start_bp"], circ=0,
                 next=np.where(s["pos"] == F["pos"] - 1, -1, s["next"]), **keep)
            _set(o, eq, pos=0, sub_pos=0, start_bp=0, circ=0, prev=-1, next=F["next"], **keep)
            _set(o, gt, pos=s["pos"] - F["pos"], sub_pos=s["sub_pos"] - F["sub_pos"],
                 start_bp=s["start_bp"] - F["start_bp"], circ=0,
                 next=np.where(idx == F["prev"], -1, s["next"]), **keep)
        else:
            rot = F["l_cont"] - (F["pos"] + 1)
            srot = F["sub_l_cont"] - (F["sub_pos"] + F["sub_len"])
            brot = F["l_cont_bp"] - (F["start_bp"] + F["len_bp"])
            _set(o, lt, pos=rot + s["pos"], sub_pos=srot + s["sub_pos"], start_bp=brot + s["start_bp"], circ=0,
                 prev=np.where(idx == F["next"], -1, s["prev"]), **keep)
            _set(o, eq, pos=rot + s["pos"], sub_pos=srot + F["sub_pos"], start_bp=brot + F["start_bp"], circ=0,
                 prev=F["prev"], next=-1, **keep)
            _set(o, gt, pos=s["pos"] - (F["pos"] + 1), sub_pos=s["sub_pos"] - (F["sub_pos"] + F["sub_len"]),
                 start_bp=s["start_bp"] - (F["start_bp"] + F["len_bp"]), circ=0,
                 prev=np.where(s["pos"] == F["pos"] + 1, -1, s["prev"]), **keep)
    return o


def paste_contigs(s, a, b, stale=None):
    """KA:3367-3693: contig(a) followed by contig(b) (reversing as needed), or circularise when a,b
    are the two ends of one contig.  ``stale``: previous contents of the output struct, kept for
    members of contig(a) when a,b share a contig but are not its ends (reference quirk Q4)."""
    o = copy_state(s)
    A, B = _piv(s, a), _piv(s, b)
    inA = s["id_c"] == A["id_c"]
    inB = s["id_c"] == B["id_c"]
    if A["id_c"] != B["id_c"]:
        tot = dict(l_cont=A["l_cont"] + B["l_cont"], l_cont_bp=A["l_cont_bp"] + B["l_cont_bp"],
                   sub_l_cont=A["sub_l_cont"] + B["sub_l_cont"])
        if A["pos"] == 0:
            _set(o, inA, pos=A["l_cont"] - (s["pos"] + 1), sub_pos=A["sub_l_cont"] - (s["sub_pos"] + s["sub_len"]),
                 start_bp=A["l_cont_bp"] - (s["start_bp"] + s["len_bp"]), circ=0, ori=-s["ori"],
                 prev=np.where(s["pos"] == A["l_cont"] - 1, -1, s["next"]),
                 next=np.where(s["pos"] == A["pos"], b, s["prev"]), **tot)
        else:
            _set(o, inA, circ=0, next=np.where(s["pos"] == A["pos"], b, s["next"]), **tot)
        if B["pos"] == 0:
            _set(o, inB, pos=A["l_cont"] + s["pos"], sub_pos=A["sub_l_cont"] + s["sub_pos"], id_c=A["id_c"],
                 start_bp=A["l_cont_bp"] + s["start_bp"], circ=0,
                 prev=np.where(s["pos"] == B["pos"], a, s["prev"]), **tot)
        else:
            _set(o, inB, pos=A["l_cont"] + (B["l_cont"] - (s["pos"] + 1)),
                 sub_pos=A["sub_l_cont"] + (B["sub_l_cont"] - (s["sub_pos"] + s["sub_len"])), id_c=A["id_c"],
                 start_bp=A["l_cont_bp"] + (B["l_cont_bp"] - (s["start_bp"] + s["len_bp"])), circ=0, ori=-s["ori"],
                 prev=np.where(s["pos"] == B["pos"], a, s["next"]),
                 next=np.where(s["pos"] == 0, -1, s["prev"]), **tot)
    else:
        if A["pos"] == 0 and B["pos"] == A["l_cont"] - 1:
            _set(o, inA, circ=1, prev=np.where(s["pos"] == A["pos"], b, s["prev"]),
                 next=np.where(s["pos"] == A["l_cont"] - 1, a, s["next"]))
        elif A["pos"] == A["l_cont"] - 1 and B["pos"] == 0:
            _set(o, inA, circ=1, prev=np.where(s["pos"] == B["pos"], a, s["prev"]),
                 next=np.where(s["pos"] == A["l_cont"] - 1, b, s["next"]))
        else:
            if stale is None:
                raise RuntimeError("paste_contigs: stale-struct case (Q4) reached without history")
            for k in FIELDS:
                o[k][inA] = stale[k][inA]
    return o


LIST_BOUNDS = (1, 3, 5, 10, 20, 50)  # CL:417-422 (first n_insert_blocks=6 entries)


def get_bounds(s, a, b, n_bounds=6):
    """KA:2124-2270.  Returns (list_valid_insert[12], f_upstream[6], f_downstream[6]); the host
    pre-fills all three with -1 (CL:1854-1856)."""
    A, B = _piv(s, a), _piv(s, b)
    same = A["id_c"] == B["id_c"]
    pa, pb, la, lb = A["pos"], B["pos"], A["l_cont"], B["l_cont"]
    ins_is_ext = (pb == 0) or (pb == lb - 1)
    valid = [-1] * (2 * n_bounds)
    pos_up = [-1] * n_bounds
    pos_down = [-1] * n_bounds
    for i in range(n_bounds):
        if i == 0:
            if same:
                if pb < pa - 1:
                    cu, cd = pb + 1, pa
                elif pb > pa + 1:
                    cd, cu = pb - 1, pa
                else:
                    cu, cd = pa, pa
            else:
                cu, cd = pa, pa
        elif i < n_bounds - 1:
            cu = max(0, pa - LIST_BOUNDS[i - 1])
            cd = min(la - 1, pa + LIST_BOUNDS[i - 1])
        else:
            cu, cd = 0, la - 1
        if same and pb <= pa and pb >= cu:
            pos_up[i] = -1
            valid[2 * i] = -1
        else:
            pos_up[i] = cu
            if cu == 0:
                if (pa - cu == 1) or ins_is_ext:
                    valid[2 * i] = -1
                    pos_up[i] = -1
                else:
                    valid[2 * i] = 1
            else:
                valid[2 * i] = 1
        if same and ((pb >= pa and pb <= cd) or (pb == pa - 1)):
            pos_down[i] = -1
            valid[2 * i + 1] = -1
        else:
            pos_down[i] = cd
            if cd == la - 1:
                if (cd - pa == 1) or ins_is_ext:
                    valid[2 * i + 1] = -1
                    pos_down[i] = -1
                else:
                    valid[2 * i + 1] = 1
            else:
                valid[2 * i + 1] = 1
    f_up = [-1] * n_bounds
    f_down = [-1] * n_bounds
    members = np.flatnonzero(s["id_c"] == A["id_c"])
    pos_m = s["pos"][members]
    for i in range(n_bounds):
        if pos_down[i] >= 0:
            hit = members[pos_m == pos_down[i]]
            if hit.size:
                f_down[i] = int(hit[-1])
        if pos_up[i] >= 0:
            hit = members[pos_m == pos_up[i]]
            if hit.size:
                f_up[i] = int(hit[-1])
    return valid, f_up, f_down


def extract_block(s, a, cut, upstream, max_id):
    """KA:2400-2721: excise [cut..a] (upstream=1) or [a..cut] (upstream=0) as contig ``max_id+1``.
    ``cut`` < 0 -> unchanged copy."""
    o = copy_state(s)
    if cut < 0:
        return o
    A, C = _piv(s, a), _piv(s, cut)
    inc = s["id_c"] == A["id_c"]
    if upstream == 1:
        size = A["pos"] - C["pos"] + 1
        ssize = A["sub_pos"] - C["sub_pos"] + A["sub_len"]
        bsize = A["start_bp"] - C["start_bp"] + A["len_bp"]
        lo, hi = C, A
    else:
        size = C["pos"] - A["pos"] + 1
        ssize = C["sub_pos"] - A["sub_pos"] + C["sub_len"]
        bsize = C["start_bp"] - A["start_bp"] + C["len_bp"]
        lo, hi = A, C
    rest = dict(circ=A["circ"], l_cont=A["l_cont"] - size, sub_l_cont=A["sub_l_cont"] - ssize,
                l_cont_bp=A["l_cont_bp"] - bsize)
    lt = inc & (s["pos"] < lo["pos"])
    mid = inc & (s["pos"] >= lo["pos"]) & (s["pos"] <= hi["pos"])
    gt = inc & (s["pos"] > hi["pos"])
    _set(o, lt, next=np.where(s["pos"] == lo["pos"] - 1, hi["next"], s["next"]), **rest)
    _set(o, mid, pos=s["pos"] - lo["pos"], sub_pos=s["sub_pos"] - lo["sub_pos"], id_c=max_id + 1,
         start_bp=s["start_bp"] - lo["start_bp"], circ=0,
         prev=np.where(s["pos"] == lo["pos"], -1, s["prev"]),
         next=np.where(s["pos"] == hi["pos"], -1, s["next"]),
         l_cont=size, sub_l_cont=ssize, l_cont_bp=bsize)
    _set(o, gt, pos=s["pos"] - size, sub_pos=s["sub_pos"] - ssize, start_bp=s["start_bp"] - bsize,
         prev=np.where(s["pos"] == hi["pos"] + 1, lo["prev"], s["prev"]), **rest)
    return o


def insert_block(s, live, a, b, cut, valid_flag, upstream):
    """KA:2724-2976: insert the contig of ``a`` (in ``s`` = output of extract_block) right of ``b``,
    reversed when upstream; otherwise the result is a copy of the LIVE scaffold."""
    A, B = _piv(s, a), _piv(s, b)
    if not (A["id_c"] != B["id_c"] and valid_flag != -1):
        return copy_state(live)
    o = copy_state(s)
    n = len(s["pos"])
    idx = np.arange(n)
    inB = s["id_c"] == B["id_c"]
    inA = s["id_c"] == A["id_c"]
    tot = dict(l_cont=B["l_cont"] + A["l_cont"], sub_l_cont=B["sub_l_cont"] + A["sub_l_cont"],
               l_cont_bp=B["l_cont_bp"] + A["l_cont_bp"])
    lt = inB & (s["pos"] < B["pos"])
    eq = inB & (s["pos"] == B["pos"])
    gt = inB & (s["pos"] > B["pos"])
    _set(o, lt, circ=B["circ"], prev=np.where((idx == B["next"]) & (B["circ"] == 1), cut, s["prev"]), **tot)
    _set(o, eq, circ=B["circ"], ori=B["ori"], next=a, **tot)
    _set(o, gt, pos=s["pos"] + A["l_cont"], sub_pos=s["sub_pos"] + A["sub_l_cont"],
         start_bp=s["start_bp"] + A["l_cont_bp"], circ=B["circ"],
         prev=np.where(s["pos"] == B["pos"] + 1, cut, s["prev"]), **tot)
    if upstream == 0:
        _set(o, inA, pos=B["pos"] + 1 + s["pos"], sub_pos=B["sub_pos"] + B["sub_len"] + s["sub_pos"],
             id_c=B["id_c"], start_bp=B["start_bp"] + B["len_bp"] + s["start_bp"], circ=B["circ"],
             prev=np.where(s["pos"] == 0, b, s["prev"]),
             next=np.where(s["pos"] == s["l_cont"] - 1, B["next"], s["next"]), **tot)
    else:
        _set(o, inA, pos=B["pos"] + 1 + (A["l_cont"] - s["pos"] - 1),
             sub_pos=B["sub_pos"] + B["sub_len"] + (A["sub_l_cont"] - s["sub_pos"] - s["sub_len"]),
             id_c=B["id_c"], start_bp=B["start_bp"] + B["len_bp"] + (A["l_cont_bp"] - s["start_bp"] - s["len_bp"]),
             circ=B["circ"], ori=-s["ori"],
             prev=np.where(s["pos"] == s["l_cont"] - 1, b, s["next"]),
             next=np.where(s["pos"] == 0, B["next"], s["prev"]), **tot)
    return o


# --------------------------------------------------------------------------------------------
def perform_mutations(live, a, b, max_id, stale_structs=None):
    """CL:1918-1923 = 8 x pop_out_pop_in (CL:1642-1778) + transloc (CL:1780-1841) +
    insert_blocks (CL:1843-1916).  Returns (24 candidate states, list_valid_insert[12])."""
    out = [None] * 24
    P = pop_out_frag(live, a, max_id)
    max_id2 = int(P["id_c"].max())
    out[0] = copy_state(P)
    out[1] = flip_frag(live, a)
    out[2] = pop_in_frag_1(P, a, b, max_id2, 1)
    out[3] = pop_in_frag_1(P, a, b, max_id2, -1)
    out[4] = pop_in_frag_2(P, a, b, max_id2, 1)
    out[5] = pop_in_frag_2(P, a, b, max_id2, -1)
    out[6] = pop_in_frag_3(P, a, b, max_id2, 1)
    out[7] = pop_in_frag_3(P, a, b, max_id2, -1)
    mode = 0
    for up_a in (0, 1):
        T1 = split_contig(live, a, up_a, max_id)
        max_id1 = int(T1["id_c"].max())
        for up_b in (0, 1):
            T2 = split_contig(T1, b, up_b, max_id1)
            stale = None if stale_structs is None else stale_structs[8 + mode]
            out[8 + mode] = paste_contigs(T2, a, b, stale=stale)
            mode += 1
    valid, f_up, f_down = get_bounds(live, a, b)
    k = 0
    for i in range(6):
        for j in (1, 0):
            cut = f_up[i] if j == 1 else f_down[i]
            E = extract_block(live, a, cut, j, max_id)
            out[12 + k] = insert_block(E, live, a, b, cut, valid[k], j)
            k += 1
    return out, valid


def apply_family(live, a, b, op, max_id, stale_structs=None):
    """CL:2094-2151 test_copy_struct: regenerate the op's family from the live scaffold and copy
    struct #op over it.  Returns (new live state, list_valid_insert or None)."""
    if op < 8:
        P = pop_out_frag(live, a, max_id)
        max_id2 = int(P["id_c"].max())
        if op == 0:
            return copy_state(P), None
        if op == 1:
            return flip_frag(live, a), None
        fn = (pop_in_frag_1, pop_in_frag_1, pop_in_frag_2, pop_in_frag_2, pop_in_frag_3, pop_in_frag_3)[op - 2]
        return fn(P, a, b, max_id2, 1 if op % 2 == 0 else -1), None
    if op < 12:
        up_a, up_b = (op - 8) // 2, (op - 8) % 2
        T1 = split_contig(live, a, up_a, max_id)
        T2 = split_contig(T1, b, up_b, int(T1["id_c"].max()))
        stale = None if stale_structs is None else stale_structs[op]
        return paste_contigs(T2, a, b, stale=stale), None
    valid, f_up, f_down = get_bounds(live, a, b)
    k = op - 12
    i, j = k // 2, (1, 0)[k % 2]
    cut = f_up[i] if j == 1 else f_down[i]
    E = extract_block(live, a, cut, j, max_id)
    return insert_block(E, live, a, b, cut, valid[k], j), valid


def extract_uniq_mutations(live, a, b, list_valid_insert, flip_eject):
    """KA:4492-4553 (reads the list_valid_insert left by the PREVIOUS get_bounds -- quirk Q3)."""
    lst = [0, 1, 2, 3] if flip_eject == 1 else [2, 3]
    if int(live["l_cont"][b]) != 1:
        lst += [4, 5, 6, 7]
    if int(live["l_cont"][a]) != 1:
        lst += [8, 9, 10, 11]
    for i in range(12, 24):
        if list_valid_insert[i - 12] != -1:
            lst.append(i)
    return lst


def explode_genome(s, perm):
    """KA:409-426 + CL:1925-1948."""
    o = copy_state(s)
    o["pos"][:] = 0
    o["start_bp"][:] = 0
    o["sub_pos"][:] = 0
    o["id_c"][:] = np.asarray(perm, dtype=np.int32)
    o["prev"][:] = -1
    o["next"][:] = -1
    o["l_cont"][:] = 1
    o["l_cont_bp"][:] = o["len_bp"]
    o["sub_l_cont"][:] = o["sub_len"]
    return o


def renumber_contigs(s):
    """CL:2715-2881 (select_uniq_id_c KA:357-406, stable host sort by length desc CL:69-77,
    make_old_2_new_id_c KA:470-482, gl_update_pos KA:4689-4692).
    Canonical tie-break = ascending index of the contig's head (pos==0) fragment, which is what a
    sequential execution of select_uniq_id_c yields.  Returns (state, n_contigs, lengths_sorted)."""
    heads = np.flatnonzero(s["pos"] == 0)
    ids = s["id_c"][heads]
    lens = s["l_cont"][heads]
    order = np.argsort(-lens.astype(np.int64), kind="stable")
    nc = len(heads)
    old2new = {}
    for rank, k in enumerate(order):
        old2new[int(ids[k])] = rank  # later duplicates overwrite, like the kernel's scatter
    o = copy_state(s)
    lut_keys = np.array(list(old2new.keys()), dtype=np.int64)
    lut_vals = np.array(list(old2new.values()), dtype=np.int64)
    srt = np.argsort(lut_keys)
    lut_keys, lut_vals = lut_keys[srt], lut_vals[srt]
    pos_in = np.searchsorted(lut_keys, s["id_c"])
    o["id_c"] = ((nc - 1) - lut_vals[pos_in]).astype(np.int32)
    return o, nc, lens[order]

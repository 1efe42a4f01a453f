// instagraal_b200 -- scaffold move semantics as PURE per-fragment functions.
//
// The reference materialises 24 full copies of the 17-array scaffold per candidate pair by
// running ~62 kernels (cuda_lib_gl_single.py:1642-1923).  Every one of those kernels is, per
// thread, a pure function of (this fragment's fields, the fields of a few pivot fragments, a few
// scalars).  Here each kernel is restated once as such a function; a mutated field of any
// fragment under any of the 24 ops is then obtained ON THE FLY by composing them with pivots
// that were evaluated once per candidate (ig_build_descriptor), so no scaffold copy is ever
// written during scoring.  Semantics follow kernel_sparse_adapt.cu ("KA") line by line, quirks
// included; the repeat machinery (rep/activ/id_d) is inert in the reference and not carried.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define IG_HD __host__ __device__ __forceinline__
#else
#define IG_HD inline
#endif

#define IG_N_OPS 24
#define IG_N_CUT 6

struct Frag {            // one fragment's mutable fields (KA:40-58 minus the inert ones)
    int pos, sub_pos, id_c, start_bp, len_bp, sub_len, circ, prev, next, l_cont, sub_l_cont, l_cont_bp, ori;
};

// AoS record used on the device: one 64-byte line per fragment (13 fields + padding), so that a
// gather of a contact's column endpoint costs two 32-byte sectors instead of 13.
struct __attribute__((aligned(16))) FragRec { Frag f; int pad[3]; };

// SoA view of a scaffold (host-side import/export layout = the reference's 13 live arrays)
struct FragSoA {
    int *pos, *sub_pos, *id_c, *start_bp, *len_bp, *sub_len, *circ, *prev, *next, *l_cont, *sub_l_cont,
        *l_cont_bp, *ori;
};

IG_HD Frag ig_load(const FragSoA& s, int i) {
    Frag f;
    f.pos = s.pos[i]; f.sub_pos = s.sub_pos[i]; f.id_c = s.id_c[i]; f.start_bp = s.start_bp[i];
    f.len_bp = s.len_bp[i]; f.sub_len = s.sub_len[i]; f.circ = s.circ[i]; f.prev = s.prev[i];
    f.next = s.next[i]; f.l_cont = s.l_cont[i]; f.sub_l_cont = s.sub_l_cont[i];
    f.l_cont_bp = s.l_cont_bp[i]; f.ori = s.ori[i];
    return f;
}
IG_HD void ig_store(const FragSoA& s, int i, const Frag& f) {
    s.pos[i] = f.pos; s.sub_pos[i] = f.sub_pos; s.id_c[i] = f.id_c; s.start_bp[i] = f.start_bp;
    s.len_bp[i] = f.len_bp; s.sub_len[i] = f.sub_len; s.circ[i] = f.circ; s.prev[i] = f.prev;
    s.next[i] = f.next; s.l_cont[i] = f.l_cont; s.sub_l_cont[i] = f.sub_l_cont;
    s.l_cont_bp[i] = f.l_cont_bp; s.ori[i] = f.ori;
}

// ---- KA:737-1078 pop_out_frag: eject fragment `a` (pivot A = its fields in the input scaffold)
IG_HD Frag ig_pop_out(Frag f, int i, const Frag& A, int a, int max_id) {
    const int l = A.l_cont;
    if (l < 2 || f.id_c != A.id_c) return f;
    Frag o = f;
    if (f.pos == A.pos) {
        o.pos = 0; o.sub_pos = 0; o.id_c = max_id + 1; o.start_bp = 0; o.circ = 0; o.ori = 1;
        o.prev = -1; o.next = -1; o.l_cont = 1; o.sub_l_cont = f.sub_len; o.l_cont_bp = f.len_bp;
        return o;
    }
    o.l_cont = f.l_cont - 1; o.sub_l_cont = f.sub_l_cont - A.sub_len; o.l_cont_bp = f.l_cont_bp - A.len_bp;
    if (l > 2) {
        if (f.pos < A.pos) {
            o.prev = (i == A.next && A.circ == 1) ? A.prev : f.prev;
            o.next = (f.pos == A.pos - 1) ? A.next : f.next;
        } else {
            o.pos = f.pos - 1; o.sub_pos = f.sub_pos - A.sub_len; o.start_bp = f.start_bp - A.len_bp;
            o.prev = (f.pos == A.pos + 1) ? A.prev : f.prev;
            o.next = (i == A.prev && A.circ == 1) ? A.next : f.next;
        }
    } else {
        o.circ = 0; o.prev = -1; o.next = -1;
        if (f.pos > A.pos) { o.pos = f.pos - 1; o.sub_pos = f.sub_pos - A.sub_len; o.start_bp = f.start_bp - A.len_bp; }
    }
    return o;
}

// ---- KA:1081-1371 pop_in_frag_1: A immediately LEFT of B, contig cut before B.
//      f/A/B are fields in the popped scaffold P.
IG_HD Frag ig_pop_in_1(Frag f, int i, const Frag& A, int a, const Frag& B, int b, int max_id, int ori_new) {
    Frag o = f;
    if (B.circ == 0) {
        const int tl = B.l_cont - B.pos + 1, tsl = B.sub_l_cont - B.sub_pos + A.sub_len,
                  tbp = B.l_cont_bp - B.start_bp + A.len_bp;
        if (i == a) {
            o.pos = 0; o.sub_pos = 0; o.start_bp = 0; o.circ = 0; o.ori = ori_new; o.prev = -1; o.next = b;
            o.id_c = max_id + 1; o.l_cont = tl; o.sub_l_cont = tsl; o.l_cont_bp = tbp;
        } else if (f.id_c == B.id_c) {
            if (f.pos < B.pos) {
                o.circ = 0; o.next = (f.pos == B.pos - 1) ? -1 : f.next;
                o.l_cont = B.pos; o.sub_l_cont = B.sub_pos; o.l_cont_bp = B.start_bp;
            } else if (f.pos == B.pos) {
                o.pos = 1; o.sub_pos = A.sub_len; o.id_c = max_id + 1; o.start_bp = A.len_bp; o.circ = 0;
                o.ori = B.ori; o.prev = a; o.next = B.next; o.l_cont = tl; o.sub_l_cont = tsl; o.l_cont_bp = tbp;
            } else {
                o.pos = f.pos - B.pos + 1; o.sub_pos = f.sub_pos - B.sub_pos + A.sub_len; o.id_c = max_id + 1;
                o.start_bp = f.start_bp - B.start_bp + A.len_bp; o.circ = 0;
                o.l_cont = tl; o.sub_l_cont = tsl; o.l_cont_bp = tbp;
            }
        }
    } else {
        const int gl = B.l_cont + 1, gsl = B.sub_l_cont + A.sub_len, gbp = B.l_cont_bp + A.len_bp;
        if (i == a) {
            o.pos = 0; o.sub_pos = 0; o.start_bp = 0; o.circ = 0; o.ori = ori_new; o.prev = -1; o.next = b;
            o.id_c = B.id_c; o.l_cont = gl; o.sub_l_cont = gsl; o.l_cont_bp = gbp;
        } else if (f.id_c == B.id_c) {
            o.circ = 0; o.l_cont = gl; o.sub_l_cont = gsl; o.l_cont_bp = gbp;
            if (f.pos < B.pos) {
                o.pos = B.l_cont - B.pos + f.pos + 1;
                o.sub_pos = B.sub_l_cont - B.sub_pos + f.sub_pos + A.sub_len;
                o.start_bp = B.l_cont_bp - B.start_bp + f.start_bp + A.len_bp;
                o.next = (f.pos == B.pos - 1) ? -1 : f.next;
            } else if (f.pos == B.pos) {
                o.pos = 1; o.sub_pos = A.sub_len; o.start_bp = A.len_bp; o.len_bp = B.len_bp; o.sub_len = B.sub_len;
                o.ori = B.ori; o.prev = a; o.next = B.next;
            } else {
                o.pos = f.pos - B.pos + 1; o.sub_pos = f.sub_pos - B.sub_pos + A.sub_len;
                o.start_bp = f.start_bp - B.start_bp + A.len_bp;
                o.next = (i == B.prev) ? -1 : f.next;
            }
        }
    }
    return o;
}

// ---- KA:1373-1686 pop_in_frag_2: A immediately RIGHT of B, contig cut after A.
IG_HD Frag ig_pop_in_2(Frag f, int i, const Frag& A, int a, const Frag& B, int b, int max_id, int ori_new) {
    Frag o = f;
    if (B.circ == 0) {
        const int hl = B.pos + 2, hbp = B.start_bp + B.len_bp + A.len_bp, hsl = B.sub_pos + B.sub_len + A.sub_len;
        if (i == a) {
            o.pos = B.pos + 1; o.sub_pos = B.sub_pos + B.sub_len; o.id_c = B.id_c; o.start_bp = B.start_bp + B.len_bp;
            o.circ = 0; o.ori = ori_new; o.prev = b; o.next = -1; o.l_cont = hl; o.l_cont_bp = hbp; o.sub_l_cont = hsl;
        } else if (f.id_c == B.id_c) {
            o.circ = 0;
            if (f.pos < B.pos) {
                o.l_cont = hl; o.l_cont_bp = hbp; o.sub_l_cont = hsl;
            } else if (f.pos == B.pos) {
                o.ori = B.ori; o.prev = B.prev; o.next = a; o.l_cont = hl; o.l_cont_bp = hbp; o.sub_l_cont = hsl;
            } else {
                o.pos = f.pos - (B.pos + 1); o.sub_pos = f.sub_pos - (B.sub_pos + B.sub_len); o.id_c = max_id + 1;
                o.start_bp = f.start_bp - (B.start_bp + B.len_bp);
                o.prev = (f.pos == B.pos + 1) ? -1 : f.prev;
                o.l_cont = B.l_cont - (B.pos + 1); o.l_cont_bp = B.l_cont_bp - (B.start_bp + B.len_bp);
                o.sub_l_cont = B.sub_l_cont - (B.sub_pos + B.sub_len);
            }
        }
    } else {
        const int gl = B.l_cont + 1, gsl = B.sub_l_cont + A.sub_len, gbp = B.l_cont_bp + A.len_bp;
        const int rot = B.l_cont - (B.pos + 1), srot = B.sub_l_cont - (B.sub_pos + B.sub_len),
                  brot = B.l_cont_bp - (B.start_bp + B.len_bp);
        if (i == a) {
            o.pos = rot + B.pos + 1; o.sub_pos = srot + B.sub_pos + B.sub_len; o.id_c = B.id_c;
            o.start_bp = brot + B.start_bp + B.len_bp; o.circ = 0; o.ori = ori_new; o.prev = b; o.next = -1;
            o.l_cont = gl; o.sub_l_cont = gsl; o.l_cont_bp = gbp;
        } else if (f.id_c == B.id_c) {
            o.circ = 0; o.l_cont = gl; o.sub_l_cont = gsl; o.l_cont_bp = gbp;
            if (f.pos < B.pos) {
                o.pos = rot + f.pos; o.sub_pos = srot + f.sub_pos; o.start_bp = brot + f.start_bp;
                o.prev = (i == B.next) ? -1 : f.prev;
            } else if (f.pos == B.pos) {
                o.pos = rot + B.pos; o.sub_pos = srot + B.sub_pos; o.start_bp = brot + B.start_bp;
                o.len_bp = B.len_bp; o.sub_len = B.sub_len; o.prev = B.prev; o.next = a;
            } else {
                o.pos = f.pos - (B.pos + 1); o.sub_pos = f.sub_pos - (B.sub_pos + B.sub_len);
                o.start_bp = f.start_bp - (B.start_bp + B.len_bp);
                o.prev = (f.pos == B.pos + 1) ? -1 : f.prev;
            }
        }
    }
    return o;
}

// ---- KA:1688-1905 pop_in_frag_3: A immediately RIGHT of B, no cut.
IG_HD Frag ig_pop_in_3(Frag f, int i, const Frag& A, int a, const Frag& B, int b, int max_id, int ori_new) {
    Frag o = f;
    const int gl = B.l_cont + 1, gsl = B.sub_l_cont + A.sub_len, gbp = B.l_cont_bp + A.len_bp;
    if (i == a) {
        o.pos = B.pos + 1; o.sub_pos = B.sub_pos + B.sub_len; o.id_c = B.id_c; o.start_bp = B.start_bp + B.len_bp;
        o.circ = B.circ; o.ori = ori_new; o.prev = b; o.next = B.next; o.l_cont = gl; o.sub_l_cont = gsl; o.l_cont_bp = gbp;
    } else if (f.id_c == B.id_c) {
        o.circ = B.circ; o.l_cont = gl; o.sub_l_cont = gsl; o.l_cont_bp = gbp;
        if (f.pos < B.pos) {
            o.prev = (i == B.next && B.circ == 1) ? a : f.prev;
        } else if (f.pos == B.pos) {
            o.ori = B.ori; o.next = a;
        } else {
            o.pos = f.pos + 1; o.sub_pos = f.sub_pos + A.sub_len; o.start_bp = f.start_bp + A.len_bp;
            o.prev = (f.pos == B.pos + 1) ? a : f.prev;
        }
    }
    return o;
}

// ---- KA:2979-3365 split_contig at pivot F (fields in the input scaffold).
IG_HD Frag ig_split(Frag f, int i, const Frag& F, int upstream, int max_id) {
    if (F.l_cont <= 1 || f.id_c != F.id_c) return f;
    Frag o = f;
    o.circ = 0;
    if (F.circ == 0) {
        if (upstream == 1) {
            const int tl = F.l_cont - F.pos, tbp = F.l_cont_bp - F.start_bp, tsl = F.sub_l_cont - F.sub_pos;
            if (f.pos < F.pos) {
                o.next = (f.pos == F.pos - 1) ? -1 : f.next;
                o.l_cont = F.pos; o.l_cont_bp = F.start_bp; o.sub_l_cont = F.sub_pos;
            } else if (f.pos == F.pos) {
                o.pos = 0; o.sub_pos = 0; o.id_c = max_id + 1; o.start_bp = 0; o.prev = -1; o.next = F.next;
                o.l_cont = tl; o.l_cont_bp = tbp; o.sub_l_cont = tsl;
            } else {
                o.pos = f.pos - F.pos; o.sub_pos = f.sub_pos - F.sub_pos; o.id_c = max_id + 1;
                o.start_bp = f.start_bp - F.start_bp; o.l_cont = tl; o.l_cont_bp = tbp; o.sub_l_cont = tsl;
            }
        } else {
            const int hl = F.pos + 1, hbp = F.start_bp + F.len_bp, hsl = F.sub_pos + F.sub_len;
            if (f.pos < F.pos) {
                o.l_cont = hl; o.l_cont_bp = hbp; o.sub_l_cont = hsl;
            } else if (f.pos == F.pos) {
                o.prev = F.prev; o.next = -1; o.l_cont = hl; o.l_cont_bp = hbp; o.sub_l_cont = hsl;
            } else {
                o.pos = f.pos - (F.pos + 1); o.sub_pos = f.sub_pos - (F.sub_pos + F.sub_len); o.id_c = max_id + 1;
                o.start_bp = f.start_bp - (F.start_bp + F.len_bp);
                o.prev = (f.pos == F.pos + 1) ? -1 : f.prev;
                o.l_cont = F.l_cont - (F.pos + 1); o.l_cont_bp = F.l_cont_bp - (F.start_bp + F.len_bp);
                o.sub_l_cont = F.sub_l_cont - (F.sub_pos + F.sub_len);
            }
        }
    } else {
        o.l_cont = F.l_cont; o.l_cont_bp = F.l_cont_bp; o.sub_l_cont = F.sub_l_cont;
        if (upstream == 1) {
            if (f.pos < F.pos) {
                o.pos = F.l_cont - F.pos + f.pos; o.sub_pos = F.sub_l_cont - F.sub_pos + f.sub_pos;
                o.start_bp = F.l_cont_bp - F.start_bp + f.start_bp;
                o.next = (f.pos == F.pos - 1) ? -1 : f.next;
            } else if (f.pos == F.pos) {
                o.pos = 0; o.sub_pos = 0; o.start_bp = 0; o.prev = -1; o.next = F.next;
            } else {
                o.pos = f.pos - F.pos; o.sub_pos = f.sub_pos - F.sub_pos; o.start_bp = f.start_bp - F.start_bp;
                o.next = (i == F.prev) ? -1 : f.next;
            }
        } else {
            const int rot = F.l_cont - (F.pos + 1), srot = F.sub_l_cont - (F.sub_pos + F.sub_len),
                      brot = F.l_cont_bp - (F.start_bp + F.len_bp);
            if (f.pos < F.pos) {
                o.pos = rot + f.pos; o.sub_pos = srot + f.sub_pos; o.start_bp = brot + f.start_bp;
                o.prev = (i == F.next) ? -1 : f.prev;
            } else if (f.pos == F.pos) {
                o.pos = rot + f.pos; o.sub_pos = srot + F.sub_pos; o.start_bp = brot + F.start_bp;
                o.prev = F.prev; o.next = -1;
            } else {
                o.pos = f.pos - (F.pos + 1); o.sub_pos = f.sub_pos - (F.sub_pos + F.sub_len);
                o.start_bp = f.start_bp - (F.start_bp + F.len_bp);
                o.prev = (f.pos == F.pos + 1) ? -1 : f.prev;
            }
        }
    }
    return o;
}

// Does the scaffold produced by ig_split contain the label max_id+1?  (host: ga.max of the
// id_contigs array written by the kernel, CL:1806,1822)
IG_HD int ig_split_new_label(const Frag& F, int upstream) {
    if (F.l_cont <= 1 || F.circ != 0) return 0;
    return upstream == 1 ? 1 : (F.pos < F.l_cont - 1 ? 1 : 0);
}

// ---- KA:3367-3693 paste_contigs.  *written == 0 reports the reference's "nothing written" case
//      (quirk Q4: unreachable in practice because both splits leave A and B at contig ends).
IG_HD Frag ig_paste(Frag f, int i, const Frag& A, int a, const Frag& B, int b, int* written) {
    Frag o = f;
    *written = 1;
    if (A.id_c != B.id_c) {
        if (f.id_c == A.id_c) {
            o.circ = 0;
            o.l_cont = A.l_cont + B.l_cont; o.l_cont_bp = A.l_cont_bp + B.l_cont_bp; o.sub_l_cont = A.sub_l_cont + B.sub_l_cont;
            if (A.pos == 0) {
                o.pos = A.l_cont - (f.pos + 1); o.sub_pos = A.sub_l_cont - (f.sub_pos + f.sub_len);
                o.start_bp = A.l_cont_bp - (f.start_bp + f.len_bp); o.ori = -f.ori;
                o.prev = (f.pos == A.l_cont - 1) ? -1 : f.next;
                o.next = (f.pos == A.pos) ? b : f.prev;
            } else {
                o.next = (f.pos == A.pos) ? b : f.next;
            }
        } else if (f.id_c == B.id_c) {
            o.circ = 0; o.id_c = A.id_c;
            o.l_cont = A.l_cont + B.l_cont; o.l_cont_bp = A.l_cont_bp + B.l_cont_bp; o.sub_l_cont = A.sub_l_cont + B.sub_l_cont;
            if (B.pos == 0) {
                o.pos = A.l_cont + f.pos; o.sub_pos = A.sub_l_cont + f.sub_pos; o.start_bp = A.l_cont_bp + f.start_bp;
                o.prev = (f.pos == B.pos) ? a : f.prev;
            } else {
                o.pos = A.l_cont + (B.l_cont - (f.pos + 1));
                o.sub_pos = A.sub_l_cont + (B.sub_l_cont - (f.sub_pos + f.sub_len));
                o.start_bp = A.l_cont_bp + (B.l_cont_bp - (f.start_bp + f.len_bp)); o.ori = -f.ori;
                o.prev = (f.pos == B.pos) ? a : f.next;
                o.next = (f.pos == 0) ? -1 : f.prev;
            }
        }
    } else if (f.id_c == A.id_c) {
        if (A.pos == 0 && B.pos == A.l_cont - 1) {
            o.circ = 1;
            o.prev = (f.pos == A.pos) ? b : f.prev;
            o.next = (f.pos == A.l_cont - 1) ? a : f.next;
        } else if (A.pos == A.l_cont - 1 && B.pos == 0) {
            o.circ = 1;
            o.prev = (f.pos == B.pos) ? a : f.prev;
            o.next = (f.pos == A.l_cont - 1) ? b : f.next;
        } else {
            *written = 0;
        }
    }
    return o;
}

// ---- KA:2400-2721 extract_block: excise [C..A] (upstream) or [A..C] as contig max_id+1.
//      A, C = live pivots of the visited fragment and of the cut fragment (cut < 0 -> copy).
IG_HD Frag ig_extract_block(Frag f, int i, const Frag& A, const Frag& C, int cut, int upstream, int max_id) {
    if (cut < 0 || f.id_c != A.id_c) return f;
    int size, ssize, bsize;
    const Frag& lo = upstream == 1 ? C : A;
    const Frag& hi = upstream == 1 ? A : C;
    size = hi.pos - lo.pos + 1;
    ssize = hi.sub_pos - lo.sub_pos + hi.sub_len;
    bsize = hi.start_bp - lo.start_bp + hi.len_bp;
    Frag o = f;
    if (f.pos >= lo.pos && f.pos <= hi.pos) {
        o.pos = f.pos - lo.pos; o.sub_pos = f.sub_pos - lo.sub_pos; o.id_c = max_id + 1;
        o.start_bp = f.start_bp - lo.start_bp; o.circ = 0;
        o.prev = (f.pos == lo.pos) ? -1 : f.prev;
        o.next = (f.pos == hi.pos) ? -1 : f.next;
        o.l_cont = size; o.sub_l_cont = ssize; o.l_cont_bp = bsize;
    } else {
        o.circ = A.circ; o.l_cont = A.l_cont - size; o.sub_l_cont = A.sub_l_cont - ssize; o.l_cont_bp = A.l_cont_bp - bsize;
        if (f.pos < lo.pos) {
            o.next = (f.pos == lo.pos - 1) ? hi.next : f.next;
        } else {
            o.pos = f.pos - size; o.sub_pos = f.sub_pos - ssize; o.start_bp = f.start_bp - bsize;
            o.prev = (f.pos == hi.pos + 1) ? lo.prev : f.prev;
        }
    }
    return o;
}

// ---- KA:2724-2976 insert_block: f = fields after extract_block, `live` = fields in the live
//      scaffold, EA/EB = pivots of A and B after extract_block.
IG_HD Frag ig_insert_block(Frag f, const Frag& live, int i, const Frag& EA, int a, const Frag& EB, int b,
                           int cut, int valid_flag, int upstream) {
    if (!(EA.id_c != EB.id_c && valid_flag != -1)) return live;
    Frag o = f;
    const int tl = EB.l_cont + EA.l_cont, tsl = EB.sub_l_cont + EA.sub_l_cont, tbp = EB.l_cont_bp + EA.l_cont_bp;
    if (f.id_c == EB.id_c) {
        o.circ = EB.circ; o.l_cont = tl; o.sub_l_cont = tsl; o.l_cont_bp = tbp;
        if (f.pos < EB.pos) {
            o.prev = (i == EB.next && EB.circ == 1) ? cut : f.prev;
        } else if (f.pos == EB.pos) {
            o.ori = EB.ori; o.next = a;
        } else {
            o.pos = f.pos + EA.l_cont; o.sub_pos = f.sub_pos + EA.sub_l_cont; o.start_bp = f.start_bp + EA.l_cont_bp;
            o.prev = (f.pos == EB.pos + 1) ? cut : f.prev;
        }
    } else if (f.id_c == EA.id_c) {
        o.id_c = EB.id_c; o.circ = EB.circ; o.l_cont = tl; o.sub_l_cont = tsl; o.l_cont_bp = tbp;
        if (upstream == 0) {
            o.pos = EB.pos + 1 + f.pos; o.sub_pos = EB.sub_pos + EB.sub_len + f.sub_pos;
            o.start_bp = EB.start_bp + EB.len_bp + f.start_bp;
            o.prev = (f.pos == 0) ? b : f.prev;
            o.next = (f.pos == f.l_cont - 1) ? EB.next : f.next;
        } else {
            o.pos = EB.pos + 1 + (EA.l_cont - f.pos - 1);
            o.sub_pos = EB.sub_pos + EB.sub_len + (EA.sub_l_cont - f.sub_pos - f.sub_len);
            o.start_bp = EB.start_bp + EB.len_bp + (EA.l_cont_bp - f.start_bp - f.len_bp);
            o.ori = -f.ori;
            o.prev = (f.pos == f.l_cont - 1) ? b : f.next;
            o.next = (f.pos == 0) ? EB.next : f.prev;
        }
    }
    return o;
}

// =============================================================================================
// Per-candidate descriptor: every pivot the 24 ops need, evaluated once from the live scaffold.
struct IgBlockOp { Frag C, EA, EB; int cut; int valid; };
struct IgDescriptor {
    int a, b, max_id;
    Frag A, B;                 // live pivots
    Frag PA, PB; int max_id2;  // pivots after pop_out(A)             (ops 0,2..7)
    Frag T1A[2], T1B[2]; int max_id1[2];   // after split at A (upA)  (ops 8..11)
    Frag T2A[2][2], T2B[2][2];             // after split at B (upB)
    IgBlockOp blk[12];         // ops 12..23 in the reference's launch order (i, j=1 then 0)
    int valid[12];             // list_valid_insert left by get_bounds for THIS pair
    int uniq[IG_N_OPS]; int n_uniq;  // extract_uniq_mutations (uses the PREVIOUS pair's valid list)
    int cut_pos_up[IG_N_CUT], cut_pos_down[IG_N_CUT];
    int f_up[IG_N_CUT], f_down[IG_N_CUT];
};

// KA:2124-2252 get_bounds, thread-0 part: cut positions + validity flags.
IG_HD void ig_get_bounds_positions(const Frag& A, const Frag& B, int* valid, int* pos_up, int* pos_down) {
    const int bounds[IG_N_CUT] = {1, 3, 5, 10, 20, 50};  // CL:417-422
    const int same = A.id_c == B.id_c;
    const int pa = A.pos, pb = B.pos, la = A.l_cont, lb = B.l_cont;
    const int ins_is_ext = (pb == 0) || (pb == lb - 1);
    for (int i = 0; i < IG_N_CUT; i++) {
        int cu, cd;
        if (i == 0) {
            if (same) {
                if (pb < pa - 1) { cu = pb + 1; cd = pa; }
                else if (pb > pa + 1) { cd = pb - 1; cu = pa; }
                else { cu = pa; cd = pa; }
            } else { cu = pa; cd = pa; }
        } else if (i < IG_N_CUT - 1) {
            cu = pa - bounds[i - 1]; if (cu < 0) cu = 0;
            cd = pa + bounds[i - 1]; if (cd > la - 1) cd = la - 1;
        } else { cu = 0; cd = la - 1; }
        if (same && pb <= pa && pb >= cu) { pos_up[i] = -1; valid[2 * i] = -1; }
        else {
            pos_up[i] = cu;
            if (cu == 0) {
                if ((pa - cu == 1) || ins_is_ext) { valid[2 * i] = -1; pos_up[i] = -1; }
                else valid[2 * i] = 1;
            } else valid[2 * i] = 1;
        }
        if (same && ((pb >= pa && pb <= cd) || (pb == pa - 1))) { pos_down[i] = -1; valid[2 * i + 1] = -1; }
        else {
            pos_down[i] = cd;
            if (cd == la - 1) {
                if ((cd - pa == 1) || ins_is_ext) { valid[2 * i + 1] = -1; pos_down[i] = -1; }
                else valid[2 * i + 1] = 1;
            } else valid[2 * i + 1] = 1;
        }
    }
}

// KA:4492-4553 extract_uniq_mutations (prev_valid = list left by the previous get_bounds, Q3).
IG_HD int ig_uniq_mutations(const Frag& A, const Frag& B, const int* prev_valid, int flip_eject, int* uniq) {
    int n = 0;
    if (flip_eject == 1) { uniq[n++] = 0; uniq[n++] = 1; }
    uniq[n++] = 2; uniq[n++] = 3;
    if (B.l_cont != 1) { uniq[n++] = 4; uniq[n++] = 5; uniq[n++] = 6; uniq[n++] = 7; }
    if (A.l_cont != 1) { uniq[n++] = 8; uniq[n++] = 9; uniq[n++] = 10; uniq[n++] = 11; }
    for (int i = 12; i < IG_N_OPS; i++) if (prev_valid[i - 12] != -1) uniq[n++] = i;
    for (int i = n; i < IG_N_OPS; i++) uniq[i] = -1;
    return n;
}

// Fill every pivot of the descriptor.  Requires d.a, d.b, d.max_id, d.f_up/f_down (cut fragment
// ids found by the parallel scan) and d.valid to be set.
template <class Loader>
IG_HD void ig_build_descriptor(IgDescriptor& d, const Loader& load) {
    const int a = d.a, b = d.b, max_id = d.max_id;
    d.A = load(a);
    d.B = load(b);
    d.PA = ig_pop_out(d.A, a, d.A, a, max_id);
    d.PB = ig_pop_out(d.B, b, d.A, a, max_id);
    d.max_id2 = max_id + (d.A.l_cont >= 2 ? 1 : 0);
    for (int ua = 0; ua < 2; ua++) {
        d.T1A[ua] = ig_split(d.A, a, d.A, ua, max_id);
        d.T1B[ua] = ig_split(d.B, b, d.A, ua, max_id);
        d.max_id1[ua] = max_id + ig_split_new_label(d.A, ua);
        for (int ub = 0; ub < 2; ub++) {
            d.T2A[ua][ub] = ig_split(d.T1A[ua], a, d.T1B[ua], ub, d.max_id1[ua]);
            d.T2B[ua][ub] = ig_split(d.T1B[ua], b, d.T1B[ua], ub, d.max_id1[ua]);
        }
    }
    int k = 0;
    for (int i = 0; i < IG_N_CUT; i++)
        for (int jj = 0; jj < 2; jj++, k++) {
            const int up = jj == 0 ? 1 : 0;
            IgBlockOp& o = d.blk[k];
            o.cut = up ? d.f_up[i] : d.f_down[i];
            o.valid = d.valid[k];
            o.C = o.cut >= 0 ? load(o.cut) : d.A;
            o.EA = ig_extract_block(d.A, a, d.A, o.C, o.cut, up, max_id);
            o.EB = ig_extract_block(d.B, b, d.A, o.C, o.cut, up, max_id);
        }
}

// Same as ig_build_descriptor, split over the lanes of one warp (lane 0..11: block ops; 12: pop
// pivots; 13, 14: the two translocation families).  Lanes >= 15 do nothing.
template <class Loader>
IG_HD void ig_build_descriptor_part(IgDescriptor& d, const Loader& load, int lane) {
    const int a = d.a, b = d.b, max_id = d.max_id;
    if (lane < 12) {
        const int k = lane, i = k >> 1, up = (k & 1) == 0 ? 1 : 0;
        IgBlockOp& o = d.blk[k];
        o.cut = up ? d.f_up[i] : d.f_down[i];
        o.valid = d.valid[k];
        o.C = o.cut >= 0 ? load(o.cut) : d.A;
        o.EA = ig_extract_block(d.A, a, d.A, o.C, o.cut, up, max_id);
        o.EB = ig_extract_block(d.B, b, d.A, o.C, o.cut, up, max_id);
    } else if (lane == 12) {
        d.PA = ig_pop_out(d.A, a, d.A, a, max_id);
        d.PB = ig_pop_out(d.B, b, d.A, a, max_id);
        d.max_id2 = max_id + (d.A.l_cont >= 2 ? 1 : 0);
    } else if (lane < 15) {
        const int ua = lane - 13;
        const Frag t1a = ig_split(d.A, a, d.A, ua, max_id);
        const Frag t1b = ig_split(d.B, b, d.A, ua, max_id);
        const int m1 = max_id + ig_split_new_label(d.A, ua);
        d.T1A[ua] = t1a; d.T1B[ua] = t1b; d.max_id1[ua] = m1;
        for (int ub = 0; ub < 2; ub++) {
            d.T2A[ua][ub] = ig_split(t1a, a, t1b, ub, m1);
            d.T2B[ua][ub] = ig_split(t1b, b, t1b, ub, m1);
        }
    }
}

// Fields of fragment i (live fields f) under op `op` of the candidate described by d.
IG_HD Frag ig_eval_op(const IgDescriptor& d, int op, const Frag& f, int i) {
    if (op == 1) { Frag o = f; if (i == d.a) o.ori = -f.ori; return o; }      // KA:612-670
    if (op < 8) {
        Frag p = ig_pop_out(f, i, d.A, d.a, d.max_id);
        if (op == 0) return p;
        const int ori_new = (op & 1) ? -1 : 1;
        if (op < 4) return ig_pop_in_1(p, i, d.PA, d.a, d.PB, d.b, d.max_id2, ori_new);
        if (op < 6) return ig_pop_in_2(p, i, d.PA, d.a, d.PB, d.b, d.max_id2, ori_new);
        return ig_pop_in_3(p, i, d.PA, d.a, d.PB, d.b, d.max_id2, ori_new);
    }
    if (op < 12) {
        const int ua = (op - 8) >> 1, ub = (op - 8) & 1;
        Frag t1 = ig_split(f, i, d.A, ua, d.max_id);
        Frag t2 = ig_split(t1, i, d.T1B[ua], ub, d.max_id1[ua]);
        int written;
        return ig_paste(t2, i, d.T2A[ua][ub], d.a, d.T2B[ua][ub], d.b, &written);
    }
    const int k = op - 12;
    const int up = (k & 1) == 0 ? 1 : 0;
    const IgBlockOp& o = d.blk[k];
    Frag e = ig_extract_block(f, i, d.A, o.C, o.cut, up, d.max_id);
    return ig_insert_block(e, f, i, o.EA, d.a, o.EB, d.b, o.cut, o.valid, up);
}

// =============================================================================================
// Rigid-motion classes.  Every op moves the fragments of the <=2 affected contigs PIECEWISE
// RIGIDLY: between two consecutive breakpoints (the positions of A, B and the 12 cut fragments)
// all fragments of a contig undergo the same motion (shift or reflection of start_bp / sub_pos,
// same target contig).  A contact whose two ends undergo the same motion keeps its distance, so
// its likelihood term cannot change mathematically -- but the reference recomputes the float32
// coordinate start_bp/1000 + offset of every shifted sub-fragment (KA:3751), so |s_i - s_j| moves
// by up to an ulp of the COORDINATE (not of s) and its terms pick up that noise.  The scoring kernel
// reads a per-candidate (row class, column class) bit table saying which mutations need evaluating:
// by default only pairs that are bit-identical are skipped (both ends at rest, or ends in two contigs
// before and after); with rigid pruning every pair that moves rigidly together is skipped too.
#define IG_MAX_BP 16      // contig(A): A, A+1, 6 upstream cuts, 6 downstream cuts (+1), B, B+1
#define IG_CLS_B0 17      // first class of contig(B) when it differs from contig(A)
#define IG_MAX_CLS 20
#define IG_BP_NONE 0x7fffffff
#define IG_CLS_SHIFT 24   // rowidx packs (class << 24) | index in the affected-row list

struct IgSig { int id_c, flip, dbp, dsp, circ; };
struct __attribute__((aligned(16))) IgMotion { int dbp, dsp, id_c, flip; };  // what the scoring kernel needs of IgSig
struct IgClassTab {
    int bp_sub[IG_MAX_BP];   // contig(A) breakpoints in sub-fragment position units (IG_BP_NONE = unused)
    int bp_sub_b[2];         // contig(B) != contig(A)
    int distinct_b, id_b;    // 1 when B lives in another contig; its label
    unsigned mask[IG_MAX_CLS * IG_MAX_CLS];  // bit u set: uniq slot u must be evaluated for this class pair
    IgMotion mot[IG_MAX_CLS * IG_N_OPS];     // [class][uniq slot]
    // Contacts of one linear contig that lie beyond d_max both in kb (s >= far_s) and in sub-fragment separation
    // (dp >= far_dp), with a margin for the largest shift any mutation applies, have the floor value v_inter
    // before AND after every mutation that does not reflect either end: bit u of farok[c1][c2] marks those.
    unsigned farok[IG_MAX_CLS * IG_MAX_CLS];
    float far_s; int far_dp;
    // Uniq slots whose two classes undergo the SAME motions give a contact between them bit-identical terms (same float32
    // operations on the same inputs): repmask = one representative per group of such slots (subset of mask),
    // members[pair][representative] = the slots of its group.  Whole-row work items of the scoring kernel evaluate the
    // representatives only and credit the result to every member.
    unsigned repmask[IG_MAX_CLS * IG_MAX_CLS];
    unsigned members[IG_MAX_CLS * IG_MAX_CLS * IG_N_OPS];
};
// do uniq slots with signatures (a1, a2) and (b1, b2) of the two classes of a pair give every contact the same term?
IG_HD int ig_class_pair_same_motion(const IgSig& a1, const IgSig& a2, const IgSig& b1, const IgSig& b2) {
    if (a1.circ | a2.circ | b1.circ | b2.circ) return 0;   // (circular contigs: the term also depends on the contig length)
    if (a1.dbp != b1.dbp || a1.dsp != b1.dsp || a1.flip != b1.flip) return 0;
    if (a2.dbp != b2.dbp || a2.dsp != b2.dsp || a2.flip != b2.flip) return 0;
    return (a1.id_c == a2.id_c) == (b1.id_c == b2.id_c);
}

// breakpoints in fragment-position units (bpf) and sub-fragment-position units (bps)
IG_HD void ig_class_breakpoints(const IgDescriptor& d, int* bpf, int* bps, int* bpbf, int* bpbs) {
    const int same = d.A.id_c == d.B.id_c;
    bpf[0] = d.A.pos;     bps[0] = d.A.sub_pos;
    bpf[1] = d.A.pos + 1; bps[1] = d.A.sub_pos + d.A.sub_len;
    for (int i = 0; i < IG_N_CUT; i++) {
        const IgBlockOp& up = d.blk[2 * i];
        const IgBlockOp& dn = d.blk[2 * i + 1];
        bpf[2 + i] = up.cut >= 0 ? up.C.pos : IG_BP_NONE;
        bps[2 + i] = up.cut >= 0 ? up.C.sub_pos : IG_BP_NONE;
        bpf[8 + i] = dn.cut >= 0 ? dn.C.pos + 1 : IG_BP_NONE;
        bps[8 + i] = dn.cut >= 0 ? dn.C.sub_pos + dn.C.sub_len : IG_BP_NONE;
    }
    bpf[14] = same ? d.B.pos : IG_BP_NONE;     bps[14] = same ? d.B.sub_pos : IG_BP_NONE;
    bpf[15] = same ? d.B.pos + 1 : IG_BP_NONE; bps[15] = same ? d.B.sub_pos + d.B.sub_len : IG_BP_NONE;
    bpbf[0] = d.B.pos;     bpbs[0] = d.B.sub_pos;
    bpbf[1] = d.B.pos + 1; bpbs[1] = d.B.sub_pos + d.B.sub_len;
}
IG_HD int ig_class_count(const int* bp, int n, int pos) {
    int c = 0;
    for (int i = 0; i < n; i++) c += (bp[i] <= pos) ? 1 : 0;
    return c;
}
// class of a (sub-)fragment at position `pos` of contig `id_c` (same units as the breakpoint lists)
IG_HD int ig_class_of(const int* bp, const int* bpb, int distinct_b, int id_b, int id_c, int pos) {
    if (distinct_b && id_c == id_b) return IG_CLS_B0 + ig_class_count(bpb, 2, pos);
    return ig_class_count(bp, IG_MAX_BP, pos);
}
// representative fragment position of class-representative `rep` (0..19), or -1 when it does not exist
IG_HD int ig_class_rep_pos(const IgDescriptor& d, const int* bpf, const int* bpbf, int rep, int* on_b) {
    const int distinct_b = d.A.id_c != d.B.id_c;
    int pos, len;
    if (rep < IG_CLS_B0) { *on_b = 0; pos = rep == 0 ? 0 : bpf[rep - 1]; len = d.A.l_cont; }
    else { if (!distinct_b) return -1; *on_b = 1; pos = rep == IG_CLS_B0 ? 0 : bpbf[rep - IG_CLS_B0 - 1]; len = d.B.l_cont; }
    if (pos == IG_BP_NONE || pos < 0 || pos >= len) return -1;
    return pos;
}
// motion of the fragments around position `pos` of contig(A) (on_b = 0) or contig(B) under op
IG_HD IgSig ig_class_signature(const IgDescriptor& d, int on_b, int pos, int op) {
    const Frag& P = on_b ? d.B : d.A;
    Frag v;
    int i;
    if (!on_b && pos == d.A.pos) { v = d.A; i = d.a; }
    else if (P.id_c == d.B.id_c && pos == d.B.pos) { v = d.B; i = d.b; }
    else {  // a virtual fragment: the motion does not depend on its own start/length
        v = P; v.pos = pos; v.sub_pos = 0; v.start_bp = 0; v.len_bp = 1; v.sub_len = 1; v.ori = 1; v.prev = -2; v.next = -2;
        i = -2;
    }
    const Frag m = ig_eval_op(d, op, v, i);
    IgSig s;
    s.id_c = m.id_c; s.flip = (m.ori != v.ori) ? 1 : 0; s.circ = m.circ;
    s.dbp = s.flip ? m.start_bp + v.start_bp + v.len_bp : m.start_bp - v.start_bp;
    s.dsp = s.flip ? m.sub_pos + v.sub_pos + v.sub_len : m.sub_pos - v.sub_pos;
    return s;
}
// far-contact shortcut (see IgClassTab::farok): no reflection and no circular contig on either side, or the
// ends land in two different contigs (inter-contig constant = the same floor value)
IG_HD int ig_class_pair_far_ok(const IgSig& s1, const IgSig& s2) {
    if (s1.id_c != s2.id_c) return 1;
    return s1.flip == 0 && s2.flip == 0 && s1.circ == 0 && s2.circ == 0;
}
// must uniq slot (signatures s1, s2 of the two classes) be evaluated for a contact between them?
//   rigid = 0: skip only what is BIT-IDENTICAL to the current state (both ends do not move at all, or the
//              ends stay in two different contigs);
//   rigid = 1: also skip pairs whose ends undergo the same shift / reflection (distance preserved
//              mathematically; the reference re-rounds the shifted float32 coordinates, this does not).
IG_HD int ig_class_pair_changed(const IgSig& s1, const IgSig& s2, int cur_same, int cur_circ, int rigid) {
    if (!cur_same) return s1.id_c == s2.id_c;   // two contigs stay two contigs: both terms are the inter-contig constant
    if (cur_circ || s1.circ || s2.circ) return 1;
    if (!(s1.id_c == s2.id_c && s1.flip == s2.flip && s1.dbp == s2.dbp && s1.dsp == s2.dsp)) return 1;
    return rigid ? 0 : !(s1.flip == 0 && s1.dbp == 0 && s1.dsp == 0);
}

// TEST-ONLY host build of the product's move semantics header (instagraal_b200/csrc/ig_moves.cuh)
// so that the per-fragment functions can be fuzzed against the oracle / the reference kernels on
// machines without a GPU.  Not part of the product library (which has no CPU path).
#include "../../instagraal_b200/csrc/ig_moves.cuh"
#include <string.h>

extern "C" int ig_host_eval_all(int n, const int* in13, int a, int b, int max_id, const int* prev_valid,
                                int flip_eject, int* out24x13, int* valid_out, int* uniq_out, int* q4_hits) {
    FragSoA s;
    int* base = const_cast<int*>(in13);
    int** fields[13] = {&s.pos, &s.sub_pos, &s.id_c, &s.start_bp, &s.len_bp, &s.sub_len, &s.circ, &s.prev,
                        &s.next, &s.l_cont, &s.sub_l_cont, &s.l_cont_bp, &s.ori};
    for (int k = 0; k < 13; k++) *fields[k] = base + (size_t)k * n;
    IgDescriptor d;
    memset(&d, 0, sizeof(d));
    d.a = a; d.b = b; d.max_id = max_id;
    Frag A = ig_load(s, a), B = ig_load(s, b);
    d.n_uniq = ig_uniq_mutations(A, B, prev_valid, flip_eject, d.uniq);
    ig_get_bounds_positions(A, B, d.valid, d.cut_pos_up, d.cut_pos_down);
    for (int i = 0; i < IG_N_CUT; i++) { d.f_up[i] = -1; d.f_down[i] = -1; }
    for (int f = 0; f < n; f++)
        if (s.id_c[f] == A.id_c)
            for (int i = 0; i < IG_N_CUT; i++) {
                if (s.pos[f] == d.cut_pos_down[i]) d.f_down[i] = f;
                if (s.pos[f] == d.cut_pos_up[i]) d.f_up[i] = f;
            }
    ig_build_descriptor(d, [&](int i) { return ig_load(s, i); });
    *q4_hits = 0;
    for (int op = 0; op < IG_N_OPS; op++) {
        FragSoA o;
        int** of[13] = {&o.pos, &o.sub_pos, &o.id_c, &o.start_bp, &o.len_bp, &o.sub_len, &o.circ, &o.prev,
                        &o.next, &o.l_cont, &o.sub_l_cont, &o.l_cont_bp, &o.ori};
        for (int k = 0; k < 13; k++) *of[k] = out24x13 + ((size_t)op * 13 + k) * n;
        for (int f = 0; f < n; f++) ig_store(o, f, ig_eval_op(d, op, ig_load(s, f), f));
    }
    for (int k = 0; k < 12; k++) valid_out[k] = d.valid[k];
    for (int k = 0; k < 24; k++) uniq_out[k] = d.uniq[k];
    return d.n_uniq;
}

// Rigid-motion classes: for every fragment of the affected contigs and every op, the motion
// signature computed from the fragment itself must equal the signature of its class computed from
// the class representative (a virtual fragment).  Returns the number of mismatches; *n_cls = number
// of distinct classes met.
extern "C" int ig_host_check_classes(int n, const int* in13, int a, int b, int max_id, int* n_cls, int* n_checked) {
    FragSoA s;
    int* base = const_cast<int*>(in13);
    int** fields[13] = {&s.pos, &s.sub_pos, &s.id_c, &s.start_bp, &s.len_bp, &s.sub_len, &s.circ, &s.prev,
                        &s.next, &s.l_cont, &s.sub_l_cont, &s.l_cont_bp, &s.ori};
    for (int k = 0; k < 13; k++) *fields[k] = base + (size_t)k * n;
    IgDescriptor d;
    memset(&d, 0, sizeof(d));
    d.a = a; d.b = b; d.max_id = max_id;
    Frag A = ig_load(s, a), B = ig_load(s, b);
    ig_get_bounds_positions(A, B, d.valid, d.cut_pos_up, d.cut_pos_down);
    for (int i = 0; i < IG_N_CUT; i++) { d.f_up[i] = -1; d.f_down[i] = -1; }
    for (int f = 0; f < n; f++)
        if (s.id_c[f] == A.id_c)
            for (int i = 0; i < IG_N_CUT; i++) {
                if (s.pos[f] == d.cut_pos_down[i]) d.f_down[i] = f;
                if (s.pos[f] == d.cut_pos_up[i]) d.f_up[i] = f;
            }
    ig_build_descriptor(d, [&](int i) { return ig_load(s, i); });
    int bpf[IG_MAX_BP], bps[IG_MAX_BP], bpbf[2], bpbs[2];
    ig_class_breakpoints(d, bpf, bps, bpbf, bpbs);
    const int distinct_b = A.id_c != B.id_c;
    IgSig sig[IG_MAX_CLS][IG_N_OPS];
    int have[IG_MAX_CLS];
    memset(have, 0, sizeof have);
    for (int rep = 0; rep < IG_MAX_CLS; rep++) {
        int on_b = 0;
        const int pos = ig_class_rep_pos(d, bpf, bpbf, rep, &on_b);
        if (pos < 0) continue;
        const int cls = on_b ? IG_CLS_B0 + ig_class_count(bpbf, 2, pos) : ig_class_count(bpf, IG_MAX_BP, pos);
        for (int op = 0; op < IG_N_OPS; op++) sig[cls][op] = ig_class_signature(d, on_b, pos, op);
        have[cls] = 1;
    }
    int bad = 0, seen[IG_MAX_CLS];
    memset(seen, 0, sizeof seen);
    *n_checked = 0;
    for (int f = 0; f < n; f++) {
        const Frag fr = ig_load(s, f);
        if (fr.id_c != A.id_c && fr.id_c != B.id_c) continue;
        const int cls = ig_class_of(bpf, bpbf, distinct_b, B.id_c, fr.id_c, fr.pos);
        const int cls_sub0 = ig_class_of(bps, bpbs, distinct_b, B.id_c, fr.id_c, fr.sub_pos);
        const int cls_sub1 = ig_class_of(bps, bpbs, distinct_b, B.id_c, fr.id_c, fr.sub_pos + fr.sub_len - 1);
        if (cls != cls_sub0 || cls != cls_sub1) bad++;  // sub-fragment units must give the same class
        if (!have[cls]) { bad++; continue; }
        seen[cls] = 1;
        for (int op = 0; op < IG_N_OPS; op++) {
            const Frag m = ig_eval_op(d, op, fr, f);
            IgSig g;
            g.id_c = m.id_c; g.flip = (m.ori != fr.ori) ? 1 : 0; g.circ = m.circ;
            g.dbp = g.flip ? m.start_bp + fr.start_bp + fr.len_bp : m.start_bp - fr.start_bp;
            g.dsp = g.flip ? m.sub_pos + fr.sub_pos + fr.sub_len : m.sub_pos - fr.sub_pos;
            const IgSig& c = sig[cls][op];
            (*n_checked)++;
            if (g.id_c != c.id_c || g.flip != c.flip || g.circ != c.circ || g.dbp != c.dbp || g.dsp != c.dsp) bad++;
        }
    }
    *n_cls = 0;
    for (int c = 0; c < IG_MAX_CLS; c++) *n_cls += seen[c];
    return bad;
}

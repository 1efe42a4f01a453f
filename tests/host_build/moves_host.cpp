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

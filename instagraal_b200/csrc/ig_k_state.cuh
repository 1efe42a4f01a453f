// instagraal_b200 -- state of the current scaffold: coordinates, full likelihood over every contact, tables.
// Part of ig_kernels.cu (included there, in this order; not a stand-alone translation unit).
#pragma once

// ------------------------------------------------------------------------------------------------
// K0: coordinates of the current scaffold (uni_fill_vect_dist, KA:3763-3822) + its zero term and
//     intra pixel count (eval_likelihood_on_zero with the CORRECT float mean, i.e. without Q1).
__global__ void __launch_bounds__(IG_THREADS)
k_coords(const FragRec* __restrict__ live, const SubRec* __restrict__ sub, CoordRec* __restrict__ coord,
         int* __restrict__ clen, int ns, const DevScalars* __restrict__ sc, float mbar, int use_test,
         double* __restrict__ part_z, int* __restrict__ part_n, int write_coords, SubX* __restrict__ subx) {
    TL(13);
    __shared__ double sm[32];
    __shared__ int sn;
    const Params p = use_test ? sc->p_test : sc->p;
    if (threadIdx.x == 0) sn = 0;
    __syncthreads();
    double z = 0.0;
    int nloc = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < ns; r += gridDim.x * blockDim.x) {
        CoordRec c; int len;
        if (write_coords) {
            SubRec s = sub[r];
            Frag f = live[s.parent].f;
            c = coords_of(f, s, &len);
            coord[r] = c; clen[r] = len;
            SubX x; x.start_bp = f.start_bp; x.len_ori = f.len_bp * f.ori; x.watson = s.watson; x.crick = s.crick;
            subx[r] = x;
        } else { c = coord[r]; len = clen[r]; }
        if (c.pos == 0) nloc += intra_pairs(len);
        z += zero_term(c.pos, len, c.s_tot, p, mbar);
    }
    if (nloc) atomicAdd(&sn, nloc);
    double tot = block_sum(z, sm);
    __syncthreads();
    if (threadIdx.x == 0) { part_z[blockIdx.x] = tot; part_n[blockIdx.x] = sn; }
}

// K1: full likelihood over every stored contact (evaluate_likelihood_sparse, KA:4374-4488).
//     A pure stream over the CSR: 8 bytes per contact + one 16-byte coordinate gather (L1/L2-resident table).
//     Warp per row; every lane takes TWO contacts per trip with one 128-bit load (rows start on any contact: the
//     trip starts at the even index below it, the out-of-row half is masked).  Per contact the term is
//         ob * log10(ex) - ex - obc(ob) + exz * log10(e)                      (KA:251-270, 4322-4353)
//     * obc(ob) (log10 ob!, two f64 log10 + sqrt for ob >= 15) depends on the observed count only: its sum over the
//       level is computed ONCE per handle (k_obc_sum) and subtracted from the first block's partial;
//     * contacts whose expectation is the floor v_inter (other contig, s outside (0, d_max)) need no transcendental:
//       their ob, count and exz are summed as integers / plain adds;
//     * the others (powf + log10) are QUEUED per warp in shared memory and evaluated 32 at a time with every lane
//       busy (near-diagonal and trans contacts alternate inside a row: evaluating in place leaves half of the lanes
//       idle through ~80 instructions); powf_pos is bit-identical to the reference's powf, log10_f32 < 4e-16.
//     Rows of circular contigs and v_inter <= 0 take the generic per-contact routine (contact_term).
struct __align__(8) LnzQ { float s; int val; };
#define IG_LNZ_QCAP 96
__device__ __forceinline__ double lnz_eval(const LnzQ e, const Params& p, double l10v) {
    const float pw = IG_POWF(e.s, p.slope);
    float exf;
    if (p.d == 2.0f) exf = __fmul_rn(__fmul_rn(p.c1, pw), p.fact);
    else exf = (p.c1 * pw * expf((p.d - 2) / (powf(e.s * p.lm / p.kuhn, 2.0f) + p.d))) * p.fact;
    exf = fmaxf(exf, p.v_inter);
    const double lg = (exf == p.v_inter) ? l10v : IG_LOG10F(exf);
    return (double)e.val * lg - (double)exf;
}
__global__ void __launch_bounds__(IG_THREADS)
k_full_lnz(const long long* __restrict__ row_ptr, const int2* __restrict__ cv, const CoordRec* __restrict__ coord,
           const int* __restrict__ clen, int ns, const DevScalars* __restrict__ sc, float mbar, int use_test,
           const float* __restrict__ exz_tab, double* __restrict__ part) {
    TL(14);
    __shared__ double sm[32];
    __shared__ LnzQ queue[IG_WARPS_PER_BLOCK][IG_LNZ_QCAP];
    const Params p = use_test ? sc->p_test : sc->p;
    const double l10v = use_test ? sc->log10_vinter_test : sc->log10_vinter;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const bool generic = !(p.v_inter > 0.0f);
    LnzQ* myq = queue[w];
    int qn = 0;                      // warp-uniform
    double acc = 0.0, acc_exz = 0.0; // queued terms / expected contacts of the zero term (floats summed in double)
    long long sob = 0;               // observed counts of the floor contacts
    int nfl = 0;                     // number of floor contacts
    for (int r = wg; r < ns; r += nw) {
        const long long b = row_ptr[r], e = row_ptr[r + 1];
        if (b == e) continue;
        const CoordRec ci = coord[r];
        if (generic || ci.s_tot != 0) {   // circular contig (rare) / degenerate floor: the generic routine, obc hoisted all the same
            const int len_i = clen[r];
            for (long long k = b + lane; k < e; k += 32) {
                const int2 c = __ldg(&cv[k]);
                const CoordRec cj = coord[c.x];
                const double ob = (double)c.y;
                // KA:4428: the circular zero term uses the ROW's contig length
                double t;
                if (generic) {   // with v_inter <= 0 a term may vanish entirely (ex == 0, KA:257): keep its own obc
                    const double obc = ob_const(ob);
                    t = contact_term(ci, cj, len_i, ob, obc, p, l10v, mbar, exz_tab) + obc;
                } else t = contact_term(ci, cj, len_i, ob, 0.0, p, l10v, mbar, exz_tab);
                acc += t;
            }
            continue;
        }
        for (long long k0 = (b & ~1LL); k0 < e; k0 += 64) {
            const long long k = k0 + 2 * lane;
            int4 c2 = make_int4(0, 0, 0, 0);
            if (k < e) c2 = __ldcs(reinterpret_cast<const int4*>(cv + k));   // streamed once: do not keep it in L1
            const bool v0 = (k >= b) && (k < e), v1 = (k + 1 < e);
            CoordRec cj0 = ci, cj1 = ci;
            if (v0) cj0 = coord[c2.x];
            if (v1) cj1 = coord[c2.z];
            bool push0 = false, push1 = false;
            float s0 = 0.f, s1 = 0.f;
            if (v0) {
                float exz = p.v_inter;
                if (cj0.id_c == ci.id_c) {
                    s0 = fabsf(ci.dist - cj0.dist);
                    exz = exz_tab[abs(ci.pos - cj0.pos)];
                    push0 = (s0 > 0.0f) && (s0 < p.d_max);
                }
                acc_exz += (double)exz;
                if (!push0) { sob += c2.y; nfl++; }
            }
            if (v1) {
                float exz = p.v_inter;
                if (cj1.id_c == ci.id_c) {
                    s1 = fabsf(ci.dist - cj1.dist);
                    exz = exz_tab[abs(ci.pos - cj1.pos)];
                    push1 = (s1 > 0.0f) && (s1 < p.d_max);
                }
                acc_exz += (double)exz;
                if (!push1) { sob += c2.w; nfl++; }
            }
            // warp-collective append (first contacts of all lanes, then second ones) + evaluation of full batches
            const unsigned m0 = __ballot_sync(0xffffffffu, push0), m1 = __ballot_sync(0xffffffffu, push1);
            if (push0) { LnzQ q; q.s = s0; q.val = c2.y; myq[qn + __popc(m0 & ((1u << lane) - 1))] = q; }
            qn += __popc(m0);
            if (push1) { LnzQ q; q.s = s1; q.val = c2.w; myq[qn + __popc(m1 & ((1u << lane) - 1))] = q; }
            qn += __popc(m1);
            __syncwarp();
            while (qn >= 32) {
                qn -= 32;
                acc += lnz_eval(myq[qn + lane], p, l10v);
            }
            __syncwarp();
        }
    }
    if (lane < qn) acc += lnz_eval(myq[lane], p, l10v);
    // floor contacts: ob * log10(v_inter) - v_inter each (KA:259-262 with ex == v_inter)
    acc += l10v * (double)sob - (double)p.v_inter * (double)nfl + acc_exz * (double)LOG10E_F;
    double tot = block_sum(acc, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = tot - (blockIdx.x == 0 ? sc->obc_total : 0.0);
}
// sum over every stored contact of the part of its term that depends on the observed count only (KA:259,262)
__global__ void __launch_bounds__(IG_THREADS)
k_obc_sum(const int2* __restrict__ cv, long long nnz, double* __restrict__ part) {
    __shared__ double sm[32];
    double acc = 0.0;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x) {
        const double ob = (double)max(cv[k].y, 0);
        acc += ob_const(ob);
    }
    const double tot = block_sum(acc, sm);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

// generic deterministic final reduction of `n` doubles (and optionally ints) by one block
__global__ void k_reduce(const double* __restrict__ part, int n, double* out, const int* __restrict__ ipart, int* iout) {
    __shared__ double sm[32];
    __shared__ int smi;
    if (threadIdx.x == 0) smi = 0;
    __syncthreads();
    double v = 0.0;
    int iv = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { v += part[i]; if (ipart) iv += ipart[i]; }
    if (ipart && iv) atomicAdd(&smi, iv);
    double tot = block_sum(v, sm);
    __syncthreads();
    if (threadIdx.x == 0) { *out = tot; if (iout) *iout = smi; }
}

// exz table: expected contacts at integer sub-fragment separation (linear contigs), KA:4330-4335
__global__ void k_exz_table(float* __restrict__ tab, int n, const DevScalars* __restrict__ sc, float mbar, int use_test) {
    const Params p = use_test ? sc->p_test : sc->p;
    for (int d = blockIdx.x * blockDim.x + threadIdx.x; d < n; d += gridDim.x * blockDim.x) {
        float s_z = __int2float_rn(d) * mbar;
        tab[d] = (s_z < p.d_max) ? rippe_contacts(s_z, p) : p.v_inter;
    }
}
__global__ void k_set_params(DevScalars* sc, Params p, int test) {
    if (test) { sc->p_test = p; sc->log10_vinter_test = log10((double)p.v_inter); }
    else { sc->p = p; sc->log10_vinter = log10((double)p.v_inter); }
}

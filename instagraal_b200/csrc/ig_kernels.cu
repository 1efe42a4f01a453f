// instagraal_b200 -- hand-written CUDA (sm_100a) for the scaffolding-MCMC hot path + its C ABI.
//
// One ig_step() = one reference step_sampler() (cuda_lib_gl_single.py:1401-1465, "CL") with the
// ~700 host<->device crossings collapsed into 11-12 stream-ordered launches (one CUDA-graph replay) and
// ONE blocking D2H of a 6 KB result record; ig_run_cycle / ig_run_cycle_device enqueue a whole sweep.
// Design (DESIGN.md has the long form):
//   * contacts: CSR by row sub-fragment of the strict upper triangle, (col,val) interleaved int2;
//     a candidate touches only the rows of the <=2 affected contigs (ordered row list built on
//     device), never the whole COO, and never through the host.
//   * the 24 candidate scaffolds are never materialised.  Every op moves the fragments between two
//     breakpoints rigidly (ig_moves.cuh): k_classes evaluates each op once per class; the row end of a
//     contact comes from a per-(row, mutation) table (k_precompute), the column end is recomputed on
//     the fly from its class motion with the reference's float32 operations.
//   * scores are sums of DIFFERENCES to the current state over the contacts whose term changes;
//     class-pair bit tables say which (contact, mutation) pairs need a look at all.
//   * two scoring paths: k_score (warp per affected row, large levels) and k_pick + k_eval_flat (packed
//     per-chunk contact lists, levels up to 1.5 M contacts).
//   * all sums in double, fixed (deterministic) reduction order: lane -> warp shuffle -> block ->
//     partial array -> single-block tree.  No floating-point atomics anywhere.
//   * expected contacts in float32 exactly as the reference kernel writes them (same libdevice
//     powf/expf/log10 calls, same association), so per-contact terms are bit-identical to the
//     reference's eval_sub_likelihood (kernel_sparse_adapt.cu:4236-4370, "KA").
// Tensor cores are deliberately unused: sparse gather + transcendental math, no contraction.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/instagraal_b200.h"
#include "ig_moves.cuh"

#ifndef IG_WARPS_PER_BLOCK
#define IG_WARPS_PER_BLOCK 8
#endif
#define IG_SCORE_CTAS_PER_SM (24 / IG_WARPS_PER_BLOCK)
#define IG_THREADS (IG_WARPS_PER_BLOCK * 32)
#define IG_ROW_CHUNK 1024
#define IG_FLAT_MAX_NNZ 1500000   // levels up to 1.5 M stored contacts take the flat scoring path (list: 64 B x nnz x 8)
#define IG_ROWS_SMALL_CHUNKS 16  // levels of up to 16 Ki sub-fragments build the affected-row list in one launch
#define IG_LANE_CUR 24  // lane that owns the current-state zero term in the score kernel

struct Params { float kuhn, lm, c1, slope, d, d_max, fact, v_inter; };  // KA:91-100
struct SubRec { int parent; float watson; float crick; int j; };          // 16 B
struct CoordRec { float dist; int id_c; int pos; float s_tot; };          // 16 B (uni_fill_vect_dist)
struct __align__(16) SubX { int start_bp; int len_ori; float watson; float crick; };  // 16 B: what a rigid motion needs to
                                                                          // recompute a sub-fragment's coordinate (len_ori = len_bp * ori)

struct __align__(16) RowInfo { int r; int cls; int n; int pad; long long b; long long seg; CoordRec ci; };  // 48 B: all
                                   // k_score needs to start on an affected row, in one round trip
struct CandInfo {  // per-candidate slice description (slice_sp_mat prologue, KA:526-551)
    int id_a, id_b, same, is_circ;
    int up_a, down_a, up_b, down_b;
    int n_rows, n_sub, row_hi, pad;
};

struct DevScalars {
    Params p;          // live parameters
    Params p_test;     // nuisance test parameters
    double log10_vinter, log10_vinter_test;
    int max_label;
    int valid[12];     // gpu_list_valid_insert
    int n_cands, a;
    int cands[IG_MAX_CANDS];
    CandInfo ci[IG_MAX_CANDS];
    double lnz_full, z_cur, lsub_cur[IG_MAX_CANDS];
    int nintra_cur;
    int win_cand, win_op;
    int q4_hits;
    int n_heads; long long sum_l_cont; long long dist_half;
    double scores[IG_MAX_CANDS * IG_N_OPS];
    double z_new[IG_MAX_CANDS * IG_N_OPS];    // zero-term sum Z of the whole scaffold under each scored move
    int nintra_new[IG_MAX_CANDS * IG_N_OPS];  // intra pixel count under each scored move
    double lnz_new[IG_MAX_CANDS * IG_N_OPS];  // full non-zero likelihood of the scaffold under each scored move
    double likelihood;
    // incremental maintenance of (coordinates, lnz_full, z_cur, nintra_cur) across steps
    double lnz_next, z_next; int nintra_next;
    int prev_k, prev_u, prev_windowed, prev_id_a, prev_n_rows;
    unsigned int ticket_out, ticket_post;
    int step_idx;
    double full_out[3];
    int full_nintra, pad_;
    unsigned int ticket_cuts[IG_MAX_CANDS], ticket_rows[IG_MAX_CANDS], ticket_fin;  // last-block-done counters
    // measurement: algorithmic traffic of the scoring kernel, accumulated over steps
    unsigned long long st_contacts, st_rows, st_frags, st_selected, st_proposals;
    // flat scoring path (small levels): per candidate, list slots owned by the affected rows (padded row lengths)
    int flat_segtotal[IG_MAX_CANDS];
    // per-candidate counters of the step (written by k_finalize): they travel with this record in ONE copy
    int res_nuniq[IG_MAX_CANDS], res_nsub[IG_MAX_CANDS];
};

struct CycleOut {  // compact per-step record of ig_run_cycle (128 B)
    double likelihood, lnz_full;
    long long dist_half, sum_l_cont;
    int n_heads, win_cand, win_op, q4_hits;
    int n_uniq[IG_MAX_CANDS], n_sub[IG_MAX_CANDS];
    int pad[8];
};

// ------------------------------------------------------------------------------------------------
// Optional on-device timeline (build with -DIG_TIMELINE, scripts/gpu_timeline.sh): every kernel of the step
// records the earliest block start and the latest block end in %globaltimer nanoseconds, per step of a cycle
// run -- the only way to see the real kernel durations AND the gaps between dependent launches inside a CUDA
// graph replay with warm caches (ncu serialises and flushes; nsys is not available here).
#define IG_TL_KERNELS 16
#define IG_TL_STEPS 4096
#ifdef IG_TIMELINE
__device__ unsigned long long g_tl[IG_TL_STEPS][IG_TL_KERNELS][2];
__device__ int g_tl_step;
struct TlScope {
    int id;
    __device__ __forceinline__ TlScope(int i) : id(i) {
        if (threadIdx.x == 0) {
            unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            atomicMin(&g_tl[*(volatile int*)&g_tl_step & (IG_TL_STEPS - 1)][id][0], t);
        }
    }
    __device__ __forceinline__ ~TlScope() {
        if (threadIdx.x == 0) {
            unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            atomicMax(&g_tl[*(volatile int*)&g_tl_step & (IG_TL_STEPS - 1)][id][1], t);
        }
    }
};
#define TL(id) TlScope tl_scope_(id)
// per-block trace of the scoring kernel (last launch wins): start, end, SM id, items processed
#define IG_TL_BLOCKS 8192
__device__ unsigned long long g_tlb[IG_TL_BLOCKS][4];
struct TlBlock {
    int idx; unsigned long long t0; int items;
    __device__ __forceinline__ TlBlock() : items(0) {
        idx = blockIdx.y * gridDim.x + blockIdx.x;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    }
    __device__ __forceinline__ ~TlBlock() {
        if (threadIdx.x == 0 && idx < IG_TL_BLOCKS) {
            unsigned long long t1; unsigned sm;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
            g_tlb[idx][0] = t0; g_tlb[idx][1] = t1; g_tlb[idx][2] = sm; g_tlb[idx][3] = (unsigned long long)items;
        }
    }
};
#define TLB() TlBlock tl_block_
#define TLB_ITEM() tl_block_.items++
// phase profile of the scoring kernel: cycles of warp 0 of every block, summed per phase
__device__ unsigned long long g_tlp[16];
#define TLP_DECL() long long tlp_t_ = clock64()
#define TLP(ph) do { if (threadIdx.x == 0) { const long long n_ = clock64(); atomicAdd(&g_tlp[ph], (unsigned long long)(n_ - tlp_t_)); tlp_t_ = n_; } } while (0)
#else
#define TLB()
#define TLB_ITEM()
#define TLP_DECL()
#define TLP(ph)
#define TL(id)
#endif

// ------------------------------------------------------------------------------------------------
// device math: textual twins of KA:111-124, 153-163, 200-225, 251-270
__device__ __forceinline__ float rippe_contacts(float s, const Params& p) {
    float result = 0.0f;
    if ((s > 0.0f) && (s < p.d_max)) {
        if (p.d == 2.0f)  // exp(0/(x+2)) == 1.0f exactly: skipping it is bit-identical
            result = (p.c1 * powf(s, p.slope)) * p.fact;
        else
            result = (p.c1 * powf(s, p.slope) * expf((p.d - 2) / (powf(s * p.lm / p.kuhn, 2.0f) + p.d))) * p.fact;
    }
    return fmaxf(result, p.v_inter);
}
__device__ __forceinline__ float rippe_contacts_circ(float s, float s_tot, const Params& p) {
    float result = 0.0f;
    if ((s > 0.0f) && (s < p.d_max)) {
        float K = p.lm / p.kuhn;
        float n = K * s * (s_tot - s) / s_tot;
        result = (powf(p.kuhn, -3.0f) * powf(n, p.slope) * expf((p.d - 2.0f) / (powf(n, 2.0f) + p.d))) * p.fact;
    }
    return fmaxf(result, p.d_max);  // sic: floored at d_max (quirk Q6, KA:219)
}
__device__ float factorial_ref(float n) {
    float result = 1;
    n = floorf(n);
    if (n < 10) { for (int c = 1; c <= n; c++) result = result * c; }
    else result = powf(n, n) * expf(-n) * sqrtf(2 * M_PI * n);
    return result;
}
__constant__ double c_log10_fact[16];
__global__ void k_init_tables(double* out16) {
    int t = threadIdx.x;
    if (t < 16) out16[t] = t == 0 ? 0.0 : log10((double)factorial_ref((float)t));
}
// part of the per-contact term that depends on the observed count only (KA:259,262)
__device__ __forceinline__ double ob_const(double ob) {
    if (ob >= 15.0) return ob * log10(ob) - ob + log10(sqrt(ob * 2.0 * M_PI));
    return c_log10_fact[(int)ob];
}
// evaluate_likelihood_pxl_double (KA:251-270) with the ob-only part hoisted
__device__ __forceinline__ double pxl_term(float exf, double ob, double obc, double log10_vinter, float v_inter) {
    double ex = (double)exf;
    if (ex == 0) return 0.0;
    double lg = (exf == v_inter) ? log10_vinter : log10(ex);
    return ob * lg - ex - obc;
}
#define LOG10E_F 0.43429448190325182f

__device__ __forceinline__ CoordRec coords_of(const Frag& f, const SubRec& s, int* len_out) {
    CoordRec c;
    const bool fw = f.ori == 1;
    c.dist = __int2float_rn(f.start_bp) / 1000.0f + (fw ? s.watson : s.crick);  // KA:3751
    c.id_c = f.id_c;
    int st = (int)(__int2float_rn(f.circ) * __int2float_rn(f.l_cont_bp) / 1000.0f);  // int local, KA:3715,3739
    c.s_tot = (float)st;
    c.pos = f.sub_pos + (fw ? s.j : f.sub_len - (s.j + 1));  // KA:3745-3749
    *len_out = f.sub_l_cont;
    return c;
}

// one contact's term for one scaffold state (KA:4322-4353)
__device__ __forceinline__ double contact_term(const CoordRec& ci, const CoordRec& cj, int len_j, double ob, double obc,
                                               const Params& p, double l10v, float mbar, const float* __restrict__ exz_tab) {
    float exf, exzf;
    if (ci.id_c == cj.id_c) {
        float s = fabsf(ci.dist - cj.dist);
        int dp = abs(ci.pos - cj.pos);
        if (ci.s_tot == 0) {
            exf = rippe_contacts(s, p);
            exzf = exz_tab[dp];
        } else {
            exf = rippe_contacts_circ(s, ci.s_tot, p);
            float s_z = __int2float_rn(dp) * mbar;
            if (s_z < p.d_max) exzf = rippe_contacts_circ(s_z, __int2float_rn(len_j) * mbar, p);
            else exzf = p.v_inter;
        }
    } else { exf = p.v_inter; exzf = p.v_inter; }
    return pxl_term(exf, ob, obc, l10v, p.v_inter) + (double)exzf * LOG10E_F;
}

// zero-term of one sub-fragment (KA:3955-3972); returns contribution to Z (<= 0)
__device__ __forceinline__ double zero_term(int pos, int len, float s_tot, const Params& p, float mbar) {
    if (pos <= 0) return 0.0;
    float s = __int2float_rn(pos) * mbar;
    double ex;
    if (s < p.d_max) {
        if (s_tot == 0) ex = (double)rippe_contacts(s, p);
        else ex = (double)rippe_contacts_circ(s, __int2float_rn(len) * mbar, p);
    } else ex = (double)p.v_inter;
    return -(ex * __int2double_rn(len - pos));
}
__device__ __forceinline__ int intra_pairs(int len) {  // int32 wrap + C division, KA:3950-3953
    int t = (int)((unsigned)len * (unsigned)(len - 1));
    return t / 2;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
// deterministic block sum of one double per thread (fixed tree); result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* sm /* >= 32 */) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = lane < (blockDim.x >> 5) ? sm[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

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
//     Warp per CSR row; lanes stride the row with coalesced 8-byte (col,val) loads.
__global__ void __launch_bounds__(IG_THREADS)
k_full_lnz(const long long* __restrict__ row_ptr, const int2* __restrict__ cv, const CoordRec* __restrict__ coord,
           const int* __restrict__ clen, int ns, const DevScalars* __restrict__ sc, float mbar, int use_test,
           const float* __restrict__ exz_tab, double* __restrict__ part) {
    TL(14);
    __shared__ double sm[32];
    const Params p = use_test ? sc->p_test : sc->p;
    const double l10v = use_test ? sc->log10_vinter_test : sc->log10_vinter;
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    double acc = 0.0;
    for (int r = wg; r < ns; r += nw) {
        const long long b = row_ptr[r], e = row_ptr[r + 1];
        if (b == e) continue;
        const CoordRec ci = coord[r];
        const int len_i = clen[r];
        for (long long k = b + lane; k < e; k += 32) {
            const int2 c = __ldg(&cv[k]);
            const CoordRec cj = coord[c.x];
            const double ob = (double)c.y;
            // KA:4428: the circular zero term uses the ROW's contig length
            acc += contact_term(ci, cj, len_i, ob, ob_const(ob), p, l10v, mbar, exz_tab);
        }
    }
    double tot = block_sum(acc, sm);
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

// ------------------------------------------------------------------------------------------------
// K2: per-candidate setup.  Thread 0 walks the candidates IN ORDER because extract_uniq_mutations
//     of candidate k reads the list_valid_insert left by get_bounds of candidate k-1 (quirk Q3).
__global__ void k_cand_setup(const FragRec* __restrict__ live, DevScalars* sc, IgDescriptor* desc, int n_bounds,
                             int first_flip_eject, const int* __restrict__ cyc_in) {
    TL(0);
    if (cyc_in) {  // cycle mode: this step's {n_cands, fragment, candidates} come from the uploaded cycle plan
        const int* src = cyc_in + (size_t)sc->step_idx * (2 + IG_MAX_CANDS);
        if (threadIdx.x < 2 + IG_MAX_CANDS) (&sc->n_cands)[threadIdx.x] = src[threadIdx.x];
        __syncthreads();
    }
    // one lane per candidate: pivots and get_bounds in parallel; only the uniq lists chain through the
    // previous candidate's validity list (quirk Q3), which goes through shared memory
    __shared__ int sv[IG_MAX_CANDS + 1][12];
    const int k = threadIdx.x;
    const int n = sc->n_cands;
    const int a = sc->a;
    const Frag A = live[a].f;
    if (k < 12) sv[0][k] = sc->valid[k];
    Frag B = A;
    int b = a;
    if (k < n) {
        b = sc->cands[k];
        B = live[b].f;
        IgDescriptor& d = desc[k];
        d.a = a; d.b = b; d.max_id = sc->max_label;
        d.A = A; d.B = B;
        ig_get_bounds_positions(A, B, d.valid, d.cut_pos_up, d.cut_pos_down);
        for (int i = 0; i < 12; i++) sv[k + 1][i] = d.valid[i];
        for (int i = 0; i < IG_N_CUT; i++) { d.f_up[i] = -1; d.f_down[i] = -1; }
        // slice windows, KA:526-551 (sub-fragment units of the live scaffold)
        CandInfo& c = sc->ci[k];
        int pfa = A.sub_pos * (A.ori == 1) + (A.sub_pos - A.sub_len) * (A.ori == -1); if (pfa < 0) pfa = 0;
        int pfb = B.sub_pos * (B.ori == 1) + (B.sub_pos - B.sub_len) * (B.ori == -1); if (pfb < 0) pfb = 0;
        c.id_a = A.id_c; c.id_b = B.id_c; c.same = A.id_c == B.id_c; c.is_circ = A.circ;
        c.up_a = max(0, pfa - n_bounds - A.sub_len); c.down_a = min(A.sub_l_cont - 1, pfa + n_bounds + A.sub_len);
        c.up_b = max(0, pfb - B.sub_len); c.down_b = min(B.sub_l_cont - 1, pfb + B.sub_len);
        c.n_rows = 0; c.n_sub = 0; c.row_hi = -1;
        sc->ticket_cuts[k] = 0; sc->ticket_rows[k] = 0;
    }
    __syncthreads();
    if (k < n) {
        IgDescriptor& d = desc[k];
        d.n_uniq = ig_uniq_mutations(A, B, sv[k], (k == 0) ? first_flip_eject : 0, d.uniq);
    }
    if (k < 12) sc->valid[k] = sv[n][k];  // state after the last candidate's get_bounds (CL:1854-1870)
    if (k == 0) { sc->ticket_fin = 0; sc->ticket_post = 0; }
}
// K3: cut fragments of get_bounds (KA:2255-2269), all candidates at once; the LAST block to finish a
//     candidate then evaluates every pivot of its descriptor (one thread).
__global__ void __launch_bounds__(256)
k_find_cuts(const FragRec* __restrict__ live, int nf, DevScalars* sc, IgDescriptor* desc, IgClassTab* __restrict__ clstab) {
    TL(1);
    const int k = blockIdx.y;
    if (k >= sc->n_cands) return;
    __shared__ int is_last;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    IgDescriptor& d = desc[k];
    if (i < nf) {
        const Frag f = live[i].f;
        if (f.id_c == d.A.id_c) {
#pragma unroll
            for (int c = 0; c < IG_N_CUT; c++) {
                if (f.pos == d.cut_pos_down[c]) d.f_down[c] = i;
                if (f.pos == d.cut_pos_up[c]) d.f_up[c] = i;
            }
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&sc->ticket_cuts[k], 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x < 32) {
        __threadfence();
        ig_build_descriptor_part(desc[k], [&](int j) { return live[j].f; }, threadIdx.x);
    }
    __syncthreads();
    // breakpoints of the rigid-motion classes (ig_moves.cuh): k_rows_write classifies the rows with them
    if (threadIdx.x == 0) {
        IgClassTab& ct = clstab[k];
        int bpf[IG_MAX_BP + 2], bps[IG_MAX_BP], bpbs[2];
        ig_class_breakpoints(d, bpf, bps, bpf + IG_MAX_BP, bpbs);
        for (int j = 0; j < IG_MAX_BP; j++) ct.bp_sub[j] = bps[j];
        ct.bp_sub_b[0] = bpbs[0]; ct.bp_sub_b[1] = bpbs[1];
        ct.distinct_b = d.A.id_c != d.B.id_c; ct.id_b = d.B.id_c;
    }
}

// K4: rigid-motion classes of each candidate (ig_moves.cuh): one motion per (class, uniq slot) from a class
//     representative, then the class-pair bit table read by k_score.  One block per candidate, on the side
//     stream (only k_score needs the result).
__global__ void __launch_bounds__(IG_N_OPS * 32)
k_classes(const DevScalars* __restrict__ sc, const IgDescriptor* __restrict__ desc, IgClassTab* __restrict__ clstab, int rigid, float mbar) {
    TL(2);
    const int k = blockIdx.x;
    if (k >= sc->n_cands) return;
    __shared__ IgDescriptor d;
    __shared__ int s_bpf[IG_MAX_BP + 2], s_have[IG_MAX_CLS];
    __shared__ IgSig s_sig[IG_MAX_CLS][IG_N_OPS];
    {
        const int* src = reinterpret_cast<const int*>(desc + k);
        int* dst = reinterpret_cast<int*>(&d);
        for (int i = threadIdx.x; i < (int)(sizeof(IgDescriptor) / 4); i += blockDim.x) dst[i] = src[i];
    }
    if (threadIdx.x < IG_MAX_CLS) s_have[threadIdx.x] = 0;
    __syncthreads();
    IgClassTab& ct = clstab[k];
    if (threadIdx.x == 0) {
        int bps[IG_MAX_BP], bpbs[2];
        ig_class_breakpoints(d, s_bpf, bps, s_bpf + IG_MAX_BP, bpbs);
    }
    __syncthreads();
    const int n_uniq = d.n_uniq;
    {   // one uniq slot per warp pass (lanes = class representatives): no divergence between ops
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        int on_b = 0;
        const int pos = lane < IG_MAX_CLS ? ig_class_rep_pos(d, s_bpf, s_bpf + IG_MAX_BP, lane, &on_b) : -1;
        const int cls = pos < 0 ? -1 : (on_b ? IG_CLS_B0 + ig_class_count(s_bpf + IG_MAX_BP, 2, pos) : ig_class_count(s_bpf, IG_MAX_BP, pos));
        if (w == 0 && cls >= 0) s_have[cls] = 1;
        for (int u = w; u < n_uniq; u += (int)(blockDim.x >> 5)) {
            if (cls < 0) continue;
            const IgSig g = ig_class_signature(d, on_b, pos, d.uniq[u]);   // representatives of one class agree
            s_sig[cls][u] = g;
            IgMotion mo; mo.dbp = g.dbp; mo.dsp = g.dsp; mo.id_c = g.id_c; mo.flip = g.flip;
            ct.mot[cls * IG_N_OPS + u] = mo;
        }
    }
    __syncthreads();
    const int circ_a = d.A.circ, circ_b = d.B.circ;
    for (int t = threadIdx.x; t < IG_MAX_CLS * IG_MAX_CLS; t += blockDim.x) {
        const int c1 = t / IG_MAX_CLS, c2 = t - c1 * IG_MAX_CLS;
        unsigned m = 0xffffffu;
        if (s_have[c1] && s_have[c2]) {
            m = 0;
            const int cur_same = (c1 >= IG_CLS_B0) == (c2 >= IG_CLS_B0);
            const int cur_circ = c1 >= IG_CLS_B0 ? circ_b : circ_a;
            for (int u = 0; u < n_uniq; u++)
                if (ig_class_pair_changed(s_sig[c1][u], s_sig[c2][u], cur_same, cur_circ, rigid)) m |= 1u << u;
        }
        ct.mask[t] = m;
        unsigned fo = 0;
        if (s_have[c1] && s_have[c2])
            for (int u = 0; u < n_uniq; u++)
                if (ig_class_pair_far_ok(s_sig[c1][u], s_sig[c2][u])) fo |= 1u << u;
        ct.farok[t] = fo;
    }
    // margin: twice the largest shift of any class under any non-reflecting mutation
    if (threadIdx.x < 32) {
        int mb = 0, ms = 0;
        for (int c = threadIdx.x; c < IG_MAX_CLS; c += 32)
            if (s_have[c])
                for (int u = 0; u < n_uniq; u++)
                    if (!s_sig[c][u].flip) { mb = max(mb, abs(s_sig[c][u].dbp)); ms = max(ms, abs(s_sig[c][u].dsp)); }
        mb = __reduce_max_sync(0xffffffffu, mb);
        ms = __reduce_max_sync(0xffffffffu, ms);
        if (threadIdx.x == 0) {
            const float d_max = sc->p.d_max;
            ct.far_s = d_max + 2.0f * (__int2float_ru(mb) / 1000.0f) + 1.0f;
            ct.far_dp = (int)ceilf(d_max / mbar) + 2 * ms + 2;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K5-7: ORDERED list of the CSR rows (sub-fragments) that belong to the <=2 affected contigs.
__device__ __forceinline__ bool row_affected(const CoordRec& c, const CandInfo& ci) { return c.id_c == ci.id_a || c.id_c == ci.id_b; }
__global__ void __launch_bounds__(IG_ROW_CHUNK)
k_rows_count(const CoordRec* __restrict__ coord, int ns, DevScalars* sc, int* __restrict__ chunk_cnt, int n_chunks) {
    TL(3);
    const int k = blockIdx.y;
    if (k >= sc->n_cands) return;
    __shared__ int is_last, carry;
    __shared__ int wsum[32];
    const int r = blockIdx.x * IG_ROW_CHUNK + threadIdx.x;
    const bool f = r < ns && row_affected(coord[r], sc->ci[k]);
    const int cnt = __syncthreads_count(f);
    if (threadIdx.x == 0) {
        chunk_cnt[k * n_chunks + blockIdx.x] = cnt;
        __threadfence();
        is_last = (atomicAdd(&sc->ticket_rows[k], 1u) == gridDim.x - 1);
        carry = 0;
    }
    __syncthreads();
    if (!is_last) return;
    // the last block to finish this candidate turns the chunk counts into exclusive offsets
    __threadfence();
    volatile int* c = chunk_cnt + k * n_chunks;
    for (int base = 0; base < n_chunks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = i < n_chunks ? c[i] : 0;
        int x = v;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        if (w == 0) {
            int s2 = lane < (blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s2, o); if (lane >= o) s2 += y; }
            wsum[lane] = s2;
        }
        __syncthreads();
        const int excl = carry + (w ? wsum[w - 1] : 0) + x - v;
        if (i < n_chunks) c[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) sc->ci[k].n_rows = carry;
}
__global__ void __launch_bounds__(IG_ROW_CHUNK)
k_rows_write(const CoordRec* __restrict__ coord, int ns, const DevScalars* __restrict__ sc, const int* __restrict__ chunk_off,
             int n_chunks, int* __restrict__ rows, int* __restrict__ rowidx, int* __restrict__ row_cnt, int rows_stride,
             const IgClassTab* __restrict__ clstab, const long long* __restrict__ row_ptr, RowInfo* __restrict__ rinfo) {
    TL(4);
    const int k = blockIdx.y;
    if (k >= sc->n_cands) return;
    __shared__ int wsum[32];
    __shared__ int s_bp[IG_MAX_BP + 4];
    if (threadIdx.x < IG_MAX_BP + 4) s_bp[threadIdx.x] = reinterpret_cast<const int*>(clstab + k)[threadIdx.x];  // bp_sub, bp_sub_b, distinct_b, id_b
    const int r = blockIdx.x * IG_ROW_CHUNK + threadIdx.x;
    CoordRec cr;
    if (r < ns) cr = coord[r];
    const bool f = r < ns && row_affected(cr, sc->ci[k]);
    const unsigned b = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(b);
    __syncthreads();
    if (w == 0) {
        int s = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += y; }
        wsum[lane] = s;
    }
    __syncthreads();
    if (f) {
        const int off = chunk_off[k * n_chunks + blockIdx.x] + (w ? wsum[w - 1] : 0) + __popc(b & ((1u << lane) - 1));
        rows[(size_t)k * rows_stride + off] = r;
        const int cls = ig_class_of(s_bp, s_bp + IG_MAX_BP, s_bp[IG_MAX_BP + 2], s_bp[IG_MAX_BP + 3], cr.id_c, cr.pos);
        rowidx[(size_t)k * rows_stride + r] = off | (cls << IG_CLS_SHIFT);
        const long long b = row_ptr[r];
        RowInfo ri; ri.r = r; ri.cls = cls; ri.n = (int)(row_ptr[r + 1] - b); ri.pad = 0; ri.b = b; ri.seg = 0; ri.ci = cr;
        rinfo[(size_t)k * rows_stride + off] = ri;
        row_cnt[(size_t)k * rows_stride + off] = 0;  // k_score (block mode) accumulates into it
    }
}

// K5-7 for small levels (a handful of row chunks): count + scan + write in ONE launch, one block per candidate
// walking the chunks with a running offset (saves a dependent launch of the step's chain).
__global__ void __launch_bounds__(IG_ROW_CHUNK)
k_rows_small(const CoordRec* __restrict__ coord, int ns, DevScalars* sc, int n_chunks, int* __restrict__ rows,
             int* __restrict__ rowidx, int* __restrict__ row_cnt, int rows_stride, const IgClassTab* __restrict__ clstab,
             const long long* __restrict__ row_ptr, RowInfo* __restrict__ rinfo) {
    TL(3);
    const int k = blockIdx.x;
    if (k >= sc->n_cands) return;
    __shared__ int wsum[32], wlen[32];
    __shared__ int s_bp[IG_MAX_BP + 4];
    if (threadIdx.x < IG_MAX_BP + 4) s_bp[threadIdx.x] = reinterpret_cast<const int*>(clstab + k)[threadIdx.x];
    const CandInfo ci_k = sc->ci[k];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int base = 0, seg_base = 0;   // rows so far / stored contacts of those rows (segment offsets of the flat list)
    for (int ch = 0; ch < n_chunks; ch++) {
        const int r = ch * IG_ROW_CHUNK + threadIdx.x;
        CoordRec cr;
        if (r < ns) cr = coord[r];
        const bool f = r < ns && row_affected(cr, ci_k);
        const unsigned b = __ballot_sync(0xffffffffu, f);
        long long rb = 0;
        int rn = 0;
        if (f) { rb = row_ptr[r]; rn = (int)(row_ptr[r + 1] - rb); }
        const int rn_pad = (rn + 31) & ~31;   // rows own whole 32-contact chunks of the flat list
        int lx = rn_pad;   // inclusive warp scan of the padded row lengths
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, lx, o); if (lane >= o) lx += y; }
        __syncthreads();   // wsum / wlen of the previous chunk fully consumed
        if (lane == 0) wsum[w] = __popc(b);
        if (lane == 31) wlen[w] = lx;
        __syncthreads();
        if (w == 0) {
            int s = wsum[lane], l = wlen[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, s, o), z = __shfl_up_sync(0xffffffffu, l, o);
                if (lane >= o) { s += y; l += z; }
            }
            wsum[lane] = s; wlen[lane] = l;
        }
        __syncthreads();
        if (f) {
            const int off = base + (w ? wsum[w - 1] : 0) + __popc(b & ((1u << lane) - 1));
            rows[(size_t)k * rows_stride + off] = r;
            const int cls = ig_class_of(s_bp, s_bp + IG_MAX_BP, s_bp[IG_MAX_BP + 2], s_bp[IG_MAX_BP + 3], cr.id_c, cr.pos);
            rowidx[(size_t)k * rows_stride + r] = off | (cls << IG_CLS_SHIFT);
            RowInfo ri; ri.r = r; ri.cls = cls; ri.n = rn; ri.pad = 0; ri.b = rb; ri.ci = cr;
            ri.seg = (long long)(seg_base + (w ? wlen[w - 1] : 0) + lx - rn_pad);
            rinfo[(size_t)k * rows_stride + off] = ri;
            row_cnt[(size_t)k * rows_stride + off] = 0;
        }
        base += wsum[31]; seg_base += wlen[31];
    }
    if (threadIdx.x == 0) { sc->ci[k].n_rows = base; sc->flat_segtotal[k] = seg_base; }
}

// ------------------------------------------------------------------------------------------------
// slice_sp_mat membership of one contact (KA:557-606, incl. the precedence quirk Q11 and dat>0)
__device__ __forceinline__ bool contact_selected(const CoordRec& ci, const CoordRec& cj, int val, const CandInfo& c) {
    bool sel;
    if ((cj.id_c == ci.id_c) && c.same && (c.is_circ == 0)) {
        const int x = min(ci.pos, cj.pos), y = max(ci.pos, cj.pos);
        sel = ((x <= c.down_a) && (y >= c.up_a)) || ((y >= c.up_b) && (x <= c.down_b));
    } else {
        sel = ((!c.same) && (cj.id_c == c.id_a)) || (cj.id_c == c.id_b);
    }
    return sel && (val > 0);
}

struct RowMut { float dist; int id_c; int pos; float s_tot; };  // one sub-fragment under one mutation

// transposed partial layout: part[(k * n_slots + slot) * n_blocks + block]
#define PART_IDX(k, nslots, slot, nblocks, blk) ((((size_t)(k) * (nslots) + (slot)) * (nblocks)) + (blk))

// K8a: mutated coordinates of every affected sub-fragment under every scored mutation, evaluated
//      ONCE per (row, mutation) (replaces fill_vect_dist x24, KA:3699-3760) + the zero terms
//      (eval_all_likelihood_on_zero_1st, KA:3919-4002) restricted to the affected contigs.
//      Block (25 warps) per tile of 32 affected rows (lane = row), warp w = uniq slot w, warp 24 = the current state.
#define IG_PRE_THREADS (25 * 32)
__global__ void __launch_bounds__(IG_PRE_THREADS)
k_precompute(const CoordRec* __restrict__ coord, const int* __restrict__ clen, const FragRec* __restrict__ live,
             const SubRec* __restrict__ sub, const DevScalars* __restrict__ sc, const IgDescriptor* __restrict__ desc_g,
             const int* __restrict__ rows, int ns, RowMut* __restrict__ table, int* __restrict__ table_len, float mbar,
             double* __restrict__ part_z,  // [cand][25][gridDim.x]
             int* __restrict__ part_i)     // [cand][25][gridDim.x]
{
    TL(5);
    const int k = blockIdx.y;
    const int n_rows = sc->ci[k].n_rows;
    if (k >= sc->n_cands) return;
    if ((int)blockIdx.x * 32 >= n_rows) {   // no tile for this block
        if (threadIdx.x < 25) {
            part_z[PART_IDX(k, 25, threadIdx.x, gridDim.x, blockIdx.x)] = 0.0;
            part_i[PART_IDX(k, 25, threadIdx.x, gridDim.x, blockIdx.x)] = 0;
        }
        return;
    }
    __shared__ IgDescriptor d;
    __shared__ double red[25];   // warp w owns uniq slot w (warp 24: the current state)
    __shared__ int redi[25];
    {
        const int* src = reinterpret_cast<const int*>(desc_g + k);
        int* dst = reinterpret_cast<int*>(&d);
        for (int i = threadIdx.x; i < (int)(sizeof(IgDescriptor) / 4); i += blockDim.x) dst[i] = src[i];
    }
    if (threadIdx.x < 25) { red[threadIdx.x] = 0.0; redi[threadIdx.x] = 0; }
    __syncthreads();
    const Params p = sc->p;
    const int n_uniq = d.n_uniq;
    const int lane = threadIdx.x & 31, slot = threadIdx.x >> 5;
    const int* my_rows = rows + (size_t)k * ns;
    // a block takes tiles of 32 rows (lane = row); warp w evaluates uniq slot w for the tile, so the op is
    // warp-uniform (no divergence between the 24 move functions) and the table writes are coalesced
    if (slot < n_uniq || slot == 24) {
        for (int tile = blockIdx.x; tile * 32 < n_rows; tile += gridDim.x) {
            const int ri = tile * 32 + lane;
            double z = 0.0;
            int ia = 0;
            if (ri < n_rows) {
                const int r = my_rows[ri];
                if (slot < 24) {
                    const SubRec si = sub[r];
                    const Frag fi = live[si.parent].f;
                    const Frag fm = ig_eval_op(d, d.uniq[slot], fi, si.parent);
                    int len;
                    const CoordRec c = coords_of(fm, si, &len);
                    RowMut m; m.dist = c.dist; m.id_c = c.id_c; m.pos = c.pos; m.s_tot = c.s_tot;
                    const size_t ti = ((size_t)k * IG_N_OPS + slot) * ns + ri;
                    table[ti] = m; table_len[ti] = len;
                    if (c.pos == 0) ia = intra_pairs(len);
                    z = zero_term(c.pos, len, c.s_tot, p, mbar);
                } else {
                    const CoordRec ci = coord[r];
                    const int len = clen[r];
                    if (ci.pos == 0) ia = intra_pairs(len);
                    z = zero_term(ci.pos, len, ci.s_tot, p, mbar);
                }
            }
            z = warp_sum(z);
            ia = __reduce_add_sync(0xffffffffu, ia);
            if (lane == 0) { red[slot] += z; redi[slot] += ia; }   // tiles are visited in a fixed order
        }
    }
    __syncthreads();
    if (threadIdx.x < 25) {
        part_z[PART_IDX(k, 25, threadIdx.x, gridDim.x, blockIdx.x)] = red[threadIdx.x];
        part_i[PART_IDX(k, 25, threadIdx.x, gridDim.x, blockIdx.x)] = redi[threadIdx.x];
    }
}

// Optional (IG_PREFETCH=1) L2 prefetch of the level's arrays at the start of a step, when they fit the L2
// comfortably (yeast-scale levels): a step is a chain of a dozen short dependent kernels, each of which takes its
// first-touch misses to HBM one latency at a time when the L2 is cold.  Measured on T: +1.5 % with the L2 flushed
// between steps, -3 % when steps run back to back (warm L2, the production case) -- hence off by default.
struct PfList { const char* p[12]; unsigned long long n[12]; int cnt; };
__global__ void k_prefetch_l2(PfList L) {
    TL(12);
    for (int a = 0; a < L.cnt; a++) {
        const unsigned long long lines = (L.n[a] + 127ull) >> 7;
        for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < lines; i += (unsigned long long)gridDim.x * blockDim.x)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(L.p[a] + (i << 7)));
    }
}

// K8b: THE scoring kernel (replaces slice_sp_mat + host sort + prepare_sparse_call +
//      extract_sub_likelihood + eval_sub_likelihood).  grid = (G, n_cands).  Work item = (affected
//      CSR row, group of GS uniq slots); a warp takes one item at a time, its lanes stride the row's
//      contacts with coalesced 8-byte (col,val) loads.  The group size adapts to the amount of work
//      (GS = 24 when there are more rows than warps, down to 1 when a candidate has only a handful
//      of rows) so small assemblies still fill the 148 SMs.  Mutated coordinates of both endpoints
//      come from the table written by k_precompute (row side: warp-uniform broadcast loads; column
//      side: rowidx gather, contiguous across neighbouring contacts).
// The mutation loop is deliberately NOT unrolled and the expensive math is instantiated once: an
// unrolled 24-way body (x4 group sizes) measured 35 warps stalled on instruction fetch per issue
// (ncu "no_instruction", profiles/r1_ncu_k_score_G.txt) -- the kernel has to fit the I-cache.
// Per-thread per-slot accumulators live in shared memory ([slot][thread], conflict-free).
//
// Divergence: for most (contact, mutation) pairs the term is cheap -- bit-identical to the current
// state's term, or the constant inter-contig / out-of-range floor -- and only a few lanes of a warp
// need powf + f64 log10 (ncu: 11 of 32 lanes active on average).  Those evaluations are therefore
// QUEUED per warp in shared memory and executed 32 at a time with all lanes busy; the result is
// added to the executing lane's accumulator (only the sum over lanes matters; the order is fixed,
// hence deterministic).
struct __align__(16) QEnt { float s; int dp; unsigned mask; int val; };  // 16 B; mask bit 31: subtract
#define IG_QCAP 64
#define IG_QSUB 0x80000000u

// term of a linear-contig contact at 0 < s < d_max WITHOUT the part that depends on the observed count only
// (it cancels in t_mut - t_cur)
__device__ __noinline__ void eval_queue(const QEnt* __restrict__ q, int n, double* __restrict__ my_acc, const Params& p,
                                        double l10v, const float* __restrict__ exz_tab) {
    const int lane = threadIdx.x & 31;
    if (lane < n) {
        const QEnt e = q[lane];
        const float exf = fmaxf((p.d == 2.0f) ? (p.c1 * powf(e.s, p.slope)) * p.fact
                                              : (p.c1 * powf(e.s, p.slope) * expf((p.d - 2) / (powf(e.s * p.lm / p.kuhn, 2.0f) + p.d))) * p.fact,
                                p.v_inter);  // rippe_contacts for 0 < s < d_max (KA:153-163)
        double t = pxl_term(exf, (double)e.val, 0.0, l10v, p.v_inter) + (double)exz_tab[e.dp] * LOG10E_F;
        if (e.mask & IG_QSUB) t = -t;
        for (unsigned m = e.mask & 0xffffffu; m; m &= m - 1) my_acc[(__ffs(m) - 1) * IG_THREADS] += t;
    }
}
// (evaluating two entries per lane in batches of 64 -- two interleaved powf/log10 chains -- was tried: no gain on
//  small levels, 12 % slower on the 1 Gb workload through register pressure)

// one selected contact as the slot loop needs it (column end + current state)
struct Ctc { int pos, start_bp, len_ori; float watson, crick; int val; float cur_s; int cur_dp; int rjc; double t_cur; int flags; };
                                                                                 // flags: 1 same contig now, 2 current term deferred

// term of contact x under uniq slot u: returns false when it is bit-identical to the current state; otherwise
// `add` = t_u - t_cur for the cheap cases, or push = true (s_m, dp_m to be evaluated through the queue; add = -t_cur)
__device__ __forceinline__ bool eval_pair(const Ctc& x, int u, const RowMut a, const IgMotion* __restrict__ g_mot, float row_s_tot,
                                          const Params& p, double l10v, double inter_const, float mbar, const float* __restrict__ exz_tab,
                                          const RowMut* __restrict__ tab, const int* __restrict__ tlen, int ns,
                                          float& s_m, int& dp_m, bool& push, double& add) {
    const int4 mo4 = __ldg(reinterpret_cast<const int4*>(g_mot + (x.rjc >> IG_CLS_SHIFT) * IG_N_OPS + u));  // dbp, dsp, id_c, flip
    const bool m_same = a.id_c == mo4.z;
    const bool cur_same = x.flags & 1;
    const double ob = (double)x.val;
    double t = pxl_term(p.v_inter, ob, 0.0, l10v, p.v_inter) + inter_const;  // different contigs: both expectations are v_inter
    if (m_same) {
        if (a.s_tot != 0) {  // circular contig (rare): mutated column end from the table, evaluated in place
            const int rj = x.rjc & ((1 << IG_CLS_SHIFT) - 1);
            const RowMut bm = tab[(size_t)u * ns + rj];
            CoordRec cim, cjm;
            cim.dist = a.dist; cim.id_c = a.id_c; cim.pos = a.pos; cim.s_tot = a.s_tot;
            cjm.dist = bm.dist; cjm.id_c = bm.id_c; cjm.pos = bm.pos; cjm.s_tot = bm.s_tot;
            t = contact_term(cim, cjm, tlen[(size_t)u * ns + rj], ob, 0.0, p, l10v, mbar, exz_tab);
        } else {
            // the column end under this mutation: start_bp and sub-position follow the class motion
            const int len_j = abs(x.len_ori);
            const bool fw = (x.len_ori > 0) != (mo4.w != 0);
            const int sb = mo4.w ? mo4.x - x.start_bp - len_j : x.start_bp + mo4.x;
            const float dj = __int2float_rn(sb) / 1000.0f + (fw ? x.watson : x.crick);  // KA:3751
            const int pj = mo4.w ? mo4.y - 1 - x.pos : x.pos + mo4.y;
            s_m = fabsf(a.dist - dj);
            dp_m = abs(a.pos - pj);
            if (cur_same && row_s_tot == 0 && s_m == x.cur_s && dp_m == x.cur_dp) return false;  // bit-identical inputs
            if (!((s_m > 0.0f) && (s_m < p.d_max)))
                t = pxl_term(p.v_inter, ob, 0.0, l10v, p.v_inter) + (double)exz_tab[dp_m] * LOG10E_F;  // floor v_inter
            else { push = true; t = 0.0; }  // needs powf + log10: queue it
        }
    } else if (!cur_same) return false;  // two contigs before and after
    add = t - x.t_cur;  // t_cur = 0 while deferred
    return true;
}

// warp-collective append to the warp's queue of expensive evaluations; a full batch of 32 is evaluated at once
__device__ __forceinline__ void queue_push(bool push, float s, int dp, unsigned mask, int val, QEnt* __restrict__ myq, int& qn,
                                           double* __restrict__ my_acc, const Params& p, double l10v, const float* __restrict__ exz_tab) {
    const unsigned pm = __ballot_sync(0xffffffffu, push);
    if (!pm) return;
    const int lane = threadIdx.x & 31;
    if (push) {
        QEnt en; en.s = s; en.dp = dp; en.mask = mask; en.val = val;
        myq[qn + __popc(pm & ((1u << lane) - 1))] = en;
    }
    qn += __popc(pm);
    __syncwarp();
    if (qn >= 32) {
        eval_queue(myq, 32, my_acc, p, l10v, exz_tab);
        __syncwarp();
        if (lane < qn - 32) { const QEnt mv = myq[32 + lane]; myq[lane] = mv; }
        qn -= 32;
        __syncwarp();
    }
}

// The per-slot sums are DIFFERENCES to the current state: D[u] = sum over the selected contacts whose term
// changes under mutation u of (t_u - t_cur); contacts that do not change contribute nothing and are not
// evaluated at all (score[u] = Lnz_full(cur) + Lz[u] + D[u] is algebraically KA:4029-4046; the part of a term
// that depends on the observed count only cancels and is left out).
//   * Which (contact, mutation) pairs need a look at all is read from the candidate's class-pair bit table.
//   * The mutated coordinate of the COLUMN end is recomputed on the fly from its current start_bp / offsets and
//     the rigid motion of its class under the mutation (same float32 operations as fill_vect_dist, KA:3751, so
//     bit-identical to the reference's 24 coordinate copies) -- no dependent global load inside the slot loop;
//     the ROW end is warp-uniform and staged from the k_precompute table into shared memory once per item.
//   * A pair whose (same-contig flag, s, sub-fragment separation) is bit-identical to the current state is
//     skipped; the rest is either a cheap constant (other contig / outside (0, d_max)) or goes to the queue.
__global__ void __launch_bounds__(IG_THREADS, IG_SCORE_CTAS_PER_SM)
k_score(const long long* __restrict__ row_ptr, const int2* __restrict__ cv, const CoordRec* __restrict__ coord,
        const int* __restrict__ clen, const DevScalars* __restrict__ sc, const IgDescriptor* __restrict__ desc_g,
        const int* __restrict__ rows, const int* __restrict__ rowidx, int ns, int* __restrict__ row_cnt,
        const RowMut* __restrict__ table, const int* __restrict__ table_len, float mbar, const float* __restrict__ exz_tab,
        double* __restrict__ part_nz,   // [cand][25][gridDim.x]  (24 uniq slots; slot 24 unused = 0)
        int* __restrict__ part_c,       // [cand][2][gridDim.x]   (contacts selected, contacts read)
        int gs_div,                     // work-splitting knob: split a row into slot groups while rows*groups < warps/gs_div
        const IgClassTab* __restrict__ clstab, const SubX* __restrict__ subx, const RowInfo* __restrict__ rinfo,
        int sparse_div)                 // deal (contact, mutation) pairs to the lanes when fewer than 32/sparse_div lanes are busy
{
    TL(6);
    TLB();
    TLP_DECL();
    const int k = blockIdx.y;
    if (k >= sc->n_cands) return;
    extern __shared__ double acc_s[];                 // [IG_N_OPS][IG_THREADS]
    __shared__ double red[IG_WARPS_PER_BLOCK][25];
    __shared__ int redi[IG_WARPS_PER_BLOCK][2];
    __shared__ QEnt queue[IG_WARPS_PER_BLOCK][IG_QCAP];
    __shared__ RowMut s_row[IG_WARPS_PER_BLOCK][IG_N_OPS];
    __shared__ unsigned s_chg[IG_WARPS_PER_BLOCK][32];
    __shared__ int s_off[IG_WARPS_PER_BLOCK][32];
    const CandInfo ci_k = sc->ci[k];
    const int n_uniq = desc_g[k].n_uniq;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int wg = blockIdx.x * IG_WARPS_PER_BLOCK + w, nw = gridDim.x * IG_WARPS_PER_BLOCK;
    // Work distribution: one warp per (affected row, group of gs uniq slots, part of the row); gs = 24 and one
    // part when there are more rows than warps, fewer slots per item and rows cut into `parts` interleaved
    // chunk sets when a candidate has few rows (yeast-scale assemblies), so that the grid stays occupied and
    // the longest row does not set the kernel's critical path.
    // (IG_BLOCK_MODE: one BLOCK per item, kept for experiments -- measured slower at yeast scale.)
    const bool block_mode = (gs_div < 0) && ci_k.n_rows < nw / (-gs_div);
    const int div = gs_div < 0 ? -gs_div : gs_div;
    int gs = IG_N_OPS, parts = 1;
    if (block_mode) {
        const int want = (int)gridDim.x / 2;
        gs = (ci_k.n_rows >= want) ? 24 : ((ci_k.n_rows * 4 >= want) ? 6 : ((ci_k.n_rows * 8 >= want) ? 3 : 1));
    } else {
        const int want = nw / div;
        if (ci_k.n_rows < want) gs = (ci_k.n_rows * 4 >= want) ? 6 : ((ci_k.n_rows * 8 >= want) ? 3 : 1);
        if (ci_k.n_rows * (IG_N_OPS / gs) * 2 <= nw) parts = 2;
        if (ci_k.n_rows * (IG_N_OPS / gs) * 4 <= nw) parts = 4;
        if (sparse_div >> 16) { gs = (sparse_div >> 16) & 0xff; parts = (sparse_div >> 24) & 0xff; }  // experiments: forced split
    }
    const int ng = IG_N_OPS / gs;
    const int n_items = ci_k.n_rows * ng * parts;
    const int it0 = block_mode ? (int)blockIdx.x : wg, it_step = block_mode ? (int)gridDim.x : nw;
    if ((block_mode ? (int)blockIdx.x : (int)blockIdx.x * IG_WARPS_PER_BLOCK) >= n_items) {  // nothing for this block
        if (threadIdx.x < 25) part_nz[PART_IDX(k, 25, threadIdx.x, gridDim.x, blockIdx.x)] = 0.0;
        if (threadIdx.x < 2) part_c[PART_IDX(k, 2, threadIdx.x, gridDim.x, blockIdx.x)] = 0;
        return;
    }
    const Params p = sc->p;
    const double l10v = sc->log10_vinter;
    const unsigned* g_mask = clstab[k].mask;    // small per-candidate tables: read through L1
    const unsigned* g_farok = clstab[k].farok;
    const IgMotion* g_mot = clstab[k].mot;
    const float far_s = clstab[k].far_s;
    const int far_dp = clstab[k].far_dp;
    if (lane < 25) red[w][lane] = 0.0;
    if (lane < 2) redi[w][lane] = 0;
    for (int u = 0; u < IG_N_OPS; u++) acc_s[u * IG_THREADS + threadIdx.x] = 0.0;  // kept zero between items (see the item epilogue)
    __syncwarp();
    const int* my_idx = rowidx + (size_t)k * ns;
    int* my_cnt = row_cnt + (size_t)k * ns;
    const RowMut* tab = table + (size_t)k * IG_N_OPS * ns;
    const int* tlen = table_len + (size_t)k * IG_N_OPS * ns;
    double* my_acc = acc_s + threadIdx.x;
    QEnt* myq = queue[w];
    RowMut* myrow = s_row[w];
    unsigned* mychg = s_chg[w];
    int* myoff = s_off[w];
    // constant term of a contact whose endpoints lie in different contigs (KA:4348-4352) minus the ob part
    const double inter_const = (double)p.v_inter * LOG10E_F;
    TLP(0);   // block prologue
    for (int it = it0; it < n_items; it += it_step) {
        TLB_ITEM();
        const int rg = it / parts, part = it - rg * parts;
        const int ri = rg / ng, g = rg - ri * ng;
        const int q_off = block_mode ? 32 * w : 32 * part, q_step = block_mode ? 32 * IG_WARPS_PER_BLOCK : 32 * parts;
        const int u0 = g * gs;
        if (u0 >= n_uniq && g != 0) continue;
        const int u1 = min(u0 + gs, n_uniq);
        const unsigned gmask = (u1 > u0) ? (((1u << (u1 - u0)) - 1u) << u0) : 0u;
        const RowInfo info = rinfo[(size_t)k * ns + ri];
        const CoordRec ci = info.ci;
        const int cls_r = info.cls;
        const unsigned* mrow = g_mask + cls_r * IG_MAX_CLS;
        const long long b = info.b, e = info.b + info.n;
        __syncwarp();
        if (u0 + lane < u1) myrow[lane] = tab[(size_t)(u0 + lane) * ns + ri];
        __syncwarp();
        TLP(1);   // item set-up (row record, row-end table entries)
        int row_sel = 0;
        int qn = 0;  // warp-uniform queue fill
        unsigned touched = 0;  // slots (relative to u0) that received a term in this item
        // the (col, val) pair and the column's coordinates are fetched one chunk ahead
        // (only when whole rows are processed: short row parts gain nothing from it)
        const bool ahead = parts == 1;
        int2 c_nxt = make_int2(0, 0);
        CoordRec cj_nxt = ci;
        if (ahead && b + q_off + lane < e) { c_nxt = __ldg(&cv[b + q_off + lane]); cj_nxt = coord[c_nxt.x]; }
        for (long long q0 = b + q_off; q0 < e; q0 += q_step) {
            const long long q = q0 + lane;
            int2 c = c_nxt;
            CoordRec cj = cj_nxt;
            if (!ahead && q < e) { c = __ldg(&cv[q]); cj = coord[c.x]; }
            if (ahead && q + q_step < e) c_nxt = __ldg(&cv[q + q_step]);
            unsigned m = 0;
            Ctc x;
            x.pos = 0; x.start_bp = 0; x.len_ori = 0; x.watson = 0.f; x.crick = 0.f; x.val = 0; x.cur_s = 0.f; x.cur_dp = 0;
            x.rjc = 0; x.t_cur = 0.0; x.flags = 0;
            if (q < e) {
                if ((cj.id_c == ci_k.id_a || cj.id_c == ci_k.id_b) && contact_selected(ci, cj, c.y, ci_k)) {
                    row_sel++;
                    x.rjc = my_idx[c.x];
                    x.pos = cj.pos; x.val = c.y;
                    x.cur_s = fabsf(ci.dist - cj.dist);
                    x.cur_dp = abs(ci.pos - cj.pos);
                    const bool cur_same = ci.id_c == cj.id_c;
                    m = __ldg(&mrow[x.rjc >> IG_CLS_SHIFT]) & gmask;
                    // far beyond d_max before and after: every non-reflecting mutation leaves the floor term
#ifndef IG_NO_FAR
                    if (m && cur_same && ci.s_tot == 0 && x.cur_s >= far_s && x.cur_dp >= far_dp)
                        m &= ~__ldg(&g_farok[cls_r * IG_MAX_CLS + (x.rjc >> IG_CLS_SHIFT)]);
#endif
                    if (m) {  // current-state term of the lanes that have something to evaluate
                        const SubX sx = subx[c.x];
                        x.start_bp = sx.start_bp; x.len_ori = sx.len_ori; x.watson = sx.watson; x.crick = sx.crick;
                        const double ob = (double)c.y;
                        if (!cur_same) x.t_cur = pxl_term(p.v_inter, ob, 0.0, l10v, p.v_inter) + inter_const;
                        else if (ci.s_tot != 0) x.t_cur = contact_term(ci, cj, clen[c.x], ob, 0.0, p, l10v, mbar, exz_tab);  // circular (rare)
                        else if (!((x.cur_s > 0.0f) && (x.cur_s < p.d_max))) x.t_cur = pxl_term(p.v_inter, ob, 0.0, l10v, p.v_inter) + (double)exz_tab[x.cur_dp] * LOG10E_F;
                        else x.flags |= 2;  // powf + log10: goes through the queue once, with the mask of changed slots
                        x.flags |= cur_same ? 1 : 0;
                    }
                }
            }
            if (ahead && q + q_step < e) cj_nxt = coord[c_nxt.x];
            const unsigned um = __reduce_or_sync(0xffffffffu, m);
            if (!um) continue;
            unsigned chg = 0;
            // number of (contact, mutation) pairs of this chunk
            const int n_pairs = __reduce_add_sync(0xffffffffu, __popc(m));
            if (n_pairs * (sparse_div & 0xffff) > __popc(um) * 32) {
                // DENSE: most lanes take part in most mutations -> loop over the mutations, lane = contact
#pragma unroll 1
                for (unsigned uw = um; uw; uw &= uw - 1) {
                    const int u = __ffs(uw) - 1;
                    float s_m = 0.f; int dp_m = 0; bool push = false;
                    if ((m >> u) & 1u) {
                        double add;
                        if (eval_pair(x, u, myrow[u - u0], g_mot, ci.s_tot, p, l10v, inter_const, mbar, exz_tab, tab, tlen, ns, s_m, dp_m, push, add)) {
                            chg |= 1u << (u - u0);
                            my_acc[(u - u0) * IG_THREADS] += add;
                        }
                    }
                    queue_push(push, s_m, dp_m, 1u << (u - u0), x.val, myq, qn, my_acc, p, l10v, exz_tab);
                }
            } else {
                // SPARSE (long contigs: only the contacts that cross a breakpoint change): the pairs are dealt
                // densely to the lanes; the executing lane fetches the contact from its owner by shuffles
                int pre = __popc(m);   // inclusive prefix over the lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += y; }
                __syncwarp();
                mychg[lane] = 0; myoff[lane] = pre - __popc(m);
                __syncwarp();
#pragma unroll 1
                for (int base = 0; base < n_pairs; base += 32) {
                    const int pi = base + lane;
                    const bool valid = pi < n_pairs;
                    int src = 0;
                    if (valid) {  // last lane whose exclusive offset is <= pi
#pragma unroll
                        for (int stp = 16; stp > 0; stp >>= 1) if (src + stp < 32 && myoff[src + stp] <= pi) src += stp;
                    }
                    const unsigned msrc = __shfl_sync(0xffffffffu, m, src);
                    Ctc y;
                    y.pos = __shfl_sync(0xffffffffu, x.pos, src); y.start_bp = __shfl_sync(0xffffffffu, x.start_bp, src);
                    y.len_ori = __shfl_sync(0xffffffffu, x.len_ori, src); y.watson = __shfl_sync(0xffffffffu, x.watson, src);
                    y.crick = __shfl_sync(0xffffffffu, x.crick, src); y.val = __shfl_sync(0xffffffffu, x.val, src);
                    y.cur_s = __shfl_sync(0xffffffffu, x.cur_s, src); y.cur_dp = __shfl_sync(0xffffffffu, x.cur_dp, src);
                    y.rjc = __shfl_sync(0xffffffffu, x.rjc, src); y.t_cur = __shfl_sync(0xffffffffu, x.t_cur, src);
                    y.flags = __shfl_sync(0xffffffffu, x.flags, src);
                    float s_m = 0.f; int dp_m = 0; bool push = false;
                    int u = u0;
                    if (valid) {
                        u = __fns(msrc, 0, pi - myoff[src] + 1);
                        double add;
                        if (eval_pair(y, u, myrow[u - u0], g_mot, ci.s_tot, p, l10v, inter_const, mbar, exz_tab, tab, tlen, ns, s_m, dp_m, push, add)) {
                            atomicOr(&mychg[src], 1u << (u - u0));
                            my_acc[(u - u0) * IG_THREADS] += add;
                        }
                    }
                    queue_push(push, s_m, dp_m, 1u << (u - u0), y.val, myq, qn, my_acc, p, l10v, exz_tab);
                }
                __syncwarp();
                chg = mychg[lane];
            }
            // the deferred current-state terms, subtracted from every slot that changed
            queue_push((x.flags & 2) && chg, x.cur_s, x.cur_dp, chg | IG_QSUB, x.val, myq, qn, my_acc, p, l10v, exz_tab);
            touched |= __reduce_or_sync(0xffffffffu, chg);
        }
        TLP(2);   // contact loop
        if (qn > 0) { eval_queue(myq, qn, my_acc, p, l10v, exz_tab); }
        __syncwarp();
        // fixed-order accumulation into this warp's slot sums (work items are visited in a fixed order); only
        // the slots that received a term are reduced, and their accumulators are put back to zero
        for (unsigned tw = touched; tw; tw &= tw - 1) {
            const int us = __ffs(tw) - 1;
            const double v = warp_sum(my_acc[us * IG_THREADS]);
            my_acc[us * IG_THREADS] = 0.0;
            if (lane == 0) red[w][u0 + us] += v;
        }
        if (g == 0) {
            row_sel = __reduce_add_sync(0xffffffffu, row_sel);
            if (lane == 0) {
                redi[w][0] += row_sel;
                if (block_mode) { if (row_sel) atomicAdd(&my_cnt[ri], row_sel); if (w == 0) redi[w][1] += (int)(e - b); }
                else if (parts > 1) { if (row_sel) atomicAdd(&my_cnt[ri], row_sel); if (part == 0) redi[w][1] += (int)(e - b); }
                else { my_cnt[ri] = row_sel; redi[w][1] += (int)(e - b); }
            }
        }
    }
    TLP(3);   // last queue flush + slot reductions of the items
    __syncthreads();
    TLP(4);   // waiting for the slowest warp of the block
    if (threadIdx.x < 25) {
        double v = 0.0;
        for (int ww = 0; ww < IG_WARPS_PER_BLOCK; ww++) v += red[ww][threadIdx.x];
        part_nz[PART_IDX(k, 25, threadIdx.x, gridDim.x, blockIdx.x)] = v;
    }
    if (threadIdx.x < 2) {
        int iv = 0;
        for (int ww = 0; ww < IG_WARPS_PER_BLOCK; ww++) iv += redi[ww][threadIdx.x];
        part_c[PART_IDX(k, 2, threadIdx.x, gridDim.x, blockIdx.x)] = iv;
    }
}
#define IG_SCORE_SMEM (IG_N_OPS * IG_THREADS * sizeof(double))

// ------------------------------------------------------------------------------------------------
// FLAT scoring path for small levels (yeast scale: a candidate touches a few hundred rows of ~100 contacts).
// There the row-per-warp kernel above is all latency: one tiny work item per warp, each a chain of dependent
// gathers plus a partly filled evaluation queue, and half of the lanes hold contacts outside the slice.  Instead:
//   k_pick      one warp per 32-contact chunk of an affected row (rows own whole chunks of a per-candidate list,
//               offsets = running sum of the padded row lengths from k_rows_small): slice membership, class-pair
//               mask, current-state term; the contacts that need a look under at least one mutation are written
//               compactly (ballot prefix) at the start of their chunk (+ the count), so the list order is fixed;
//   k_eval_flat work item = (one chunk's packed contacts, group of uniq slots): no selection, no gathers by
//               column, no per-row set-up; loop over the group's mutations exactly like the dense schedule of
//               k_score (same eval_pair / queue code), one accumulator reduction per block.
// Results are the same sums in a different (still fixed) order.
#define IG_PICK_PARTS 4
// blocks [first, first + count) of the k_eval_flat grid that work for a candidate: one block per non-empty
// candidate + the rest in proportion to the number of chunks
// (k >= 0: the range of candidate k; k < 0: the candidate whose range holds block `blk`, -1 in *cand if none)
__device__ __forceinline__ void flat_block_range(const DevScalars* __restrict__ sc, int n_cands, int grid, int k, int blk,
                                                 int* first, int* count, int* tiles, int* cand) {
    int t[IG_MAX_CANDS];
#pragma unroll
    for (int c = 0; c < IG_MAX_CANDS; c++) t[c] = sc->flat_segtotal[c] >> 5;   // independent loads, issued together
    int tiles_all = 0, n_nonempty = 0;
#pragma unroll
    for (int c = 0; c < IG_MAX_CANDS; c++) { if (c >= n_cands) t[c] = 0; tiles_all += t[c]; n_nonempty += t[c] > 0; }
    const int spare = grid - n_nonempty;
    int b0 = 0;
    *first = 0; *count = 0; *tiles = 0; *cand = -1;
#pragma unroll
    for (int c = 0; c < IG_MAX_CANDS; c++) {
        if (t[c] == 0) continue;
        const int nb = 1 + (int)(((long long)spare * t[c]) / tiles_all);
        const bool hit = k >= 0 ? (c == k) : (blk >= b0 && blk < b0 + nb);
        if (hit) { *first = b0; *count = nb; *tiles = t[c]; *cand = c; }
        b0 += nb;
    }
}
struct __align__(16) FlatRec {   // 64 B
    int pos, start_bp, len_ori; float watson; float crick; int val; float cur_s; int cur_dp;
    int rjc, flags; double t_cur; int ri; unsigned m; int pad[2];
};                               // flags: 1 same contig now, 2 current term deferred, 4 the row's contig is circular

__global__ void __launch_bounds__(IG_THREADS, IG_SCORE_CTAS_PER_SM)
k_pick(const int2* __restrict__ cv, const CoordRec* __restrict__ coord, const int* __restrict__ clen, DevScalars* sc,
       const IgDescriptor* __restrict__ desc_g, const int* __restrict__ rowidx, int ns, int* __restrict__ row_cnt,
       int* __restrict__ flat_cnt, size_t chunk_stride, int* __restrict__ part_c, FlatRec* __restrict__ flat, size_t flat_stride,
       float mbar, const float* __restrict__ exz_tab, const IgClassTab* __restrict__ clstab, const SubX* __restrict__ subx,
       const RowInfo* __restrict__ rinfo) {
    TL(15);
    TLP_DECL();
    const int k = blockIdx.y;
    if (k >= sc->n_cands) return;
    __shared__ int s_sel, s_read;
    const CandInfo ci_k = sc->ci[k];
    const int n_items = ci_k.n_rows * IG_PICK_PARTS;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int wg = blockIdx.x * IG_WARPS_PER_BLOCK + w, nw = gridDim.x * IG_WARPS_PER_BLOCK;
    if (threadIdx.x == 0) { s_sel = 0; s_read = 0; }
    __syncthreads();
    if ((int)blockIdx.x * IG_WARPS_PER_BLOCK < n_items) {
        const Params p = sc->p;
        const double l10v = sc->log10_vinter;
        const double inter_const = (double)p.v_inter * LOG10E_F;
        const unsigned allmask = (1u << desc_g[k].n_uniq) - 1u;
        const unsigned* g_mask = clstab[k].mask;
        const unsigned* g_farok = clstab[k].farok;
        const float far_s = clstab[k].far_s;
        const int far_dp = clstab[k].far_dp;
        const int* my_idx = rowidx + (size_t)k * ns;
        FlatRec* my_flat = flat + (size_t)k * flat_stride;
        int* my_cnt = flat_cnt + (size_t)k * chunk_stride;
        TLP(8);   // prologue
        int sel_w = 0, read_w = 0;
        for (int it = wg; it < n_items; it += nw) {
            const int ri = it / IG_PICK_PARTS, part = it - ri * IG_PICK_PARTS;
            const RowInfo info = rinfo[(size_t)k * ns + ri];
            const CoordRec ci = info.ci;
            const unsigned* mrow = g_mask + info.cls * IG_MAX_CLS;
            const long long b = info.b, e = info.b + info.n;
            int row_sel = 0;
            for (long long q0 = b + 32 * part; q0 < e; q0 += 32 * IG_PICK_PARTS) {
                const long long q = q0 + lane;
                FlatRec x;
                x.m = 0;
                if (q < e) {
                    const int2 c = __ldg(&cv[q]);
                    const CoordRec cj = coord[c.x];
                    if ((cj.id_c == ci_k.id_a || cj.id_c == ci_k.id_b) && contact_selected(ci, cj, c.y, ci_k)) {
                        row_sel++;
                        x.rjc = my_idx[c.x];
                        x.pos = cj.pos; x.val = c.y; x.ri = ri;
                        x.cur_s = fabsf(ci.dist - cj.dist);
                        x.cur_dp = abs(ci.pos - cj.pos);
                        const bool cur_same = ci.id_c == cj.id_c;
                        unsigned m = __ldg(&mrow[x.rjc >> IG_CLS_SHIFT]) & allmask;
                        if (m && cur_same && ci.s_tot == 0 && x.cur_s >= far_s && x.cur_dp >= far_dp)
                            m &= ~__ldg(&g_farok[info.cls * IG_MAX_CLS + (x.rjc >> IG_CLS_SHIFT)]);
                        if (m) {
                            const SubX sx = subx[c.x];
                            x.start_bp = sx.start_bp; x.len_ori = sx.len_ori; x.watson = sx.watson; x.crick = sx.crick;
                            const double ob = (double)c.y;
                            x.flags = (cur_same ? 1 : 0) | (ci.s_tot != 0 ? 4 : 0);
                            x.t_cur = 0.0;
                            if (!cur_same) x.t_cur = pxl_term(p.v_inter, ob, 0.0, l10v, p.v_inter) + inter_const;
                            else if (ci.s_tot != 0) x.t_cur = contact_term(ci, cj, clen[c.x], ob, 0.0, p, l10v, mbar, exz_tab);
                            else if (!((x.cur_s > 0.0f) && (x.cur_s < p.d_max))) x.t_cur = pxl_term(p.v_inter, ob, 0.0, l10v, p.v_inter) + (double)exz_tab[x.cur_dp] * LOG10E_F;
                            else x.flags |= 2;
                            x.pad[0] = 0; x.pad[1] = 0;
                        }
                        x.m = m;
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, x.m != 0);
                const long long slot0 = info.seg + (q0 - b);   // this chunk's 32 list slots
                if (x.m) my_flat[slot0 + __popc(bal & ((1u << lane) - 1))] = x;
                if (lane == 0) my_cnt[slot0 >> 5] = __popc(bal);
            }
            row_sel = __reduce_add_sync(0xffffffffu, row_sel);
            if (lane == 0 && row_sel) atomicAdd(&row_cnt[(size_t)k * ns + ri], row_sel);   // zeroed by k_rows_small
            sel_w += row_sel;
            if (part == 0) read_w += info.n;
        }
        TLP(9);   // chunk loop
        if (lane == 0 && (sel_w | read_w)) { atomicAdd(&s_sel, sel_w); atomicAdd(&s_read, read_w); }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        part_c[PART_IDX(k, 2, 0, gridDim.x, blockIdx.x)] = s_sel;
        part_c[PART_IDX(k, 2, 1, gridDim.x, blockIdx.x)] = s_read;
    }
}

__global__ void __launch_bounds__(IG_THREADS, IG_SCORE_CTAS_PER_SM)
k_eval_flat(const DevScalars* __restrict__ sc, const IgDescriptor* __restrict__ desc_g, int ns, const int* __restrict__ flat_cnt,
            size_t chunk_stride, const FlatRec* __restrict__ flat, size_t flat_stride, const RowMut* __restrict__ table,
            const int* __restrict__ table_len, float mbar, const float* __restrict__ exz_tab, double* __restrict__ part_nz,
            const IgClassTab* __restrict__ clstab, int items_per_warp) {
    TL(6);
    TLP_DECL();
    extern __shared__ double acc_s[];                 // [IG_N_OPS][IG_THREADS]
    __shared__ double red[IG_WARPS_PER_BLOCK][IG_N_OPS];
    __shared__ QEnt queue[IG_WARPS_PER_BLOCK][IG_QCAP];
    // The blocks of ONE grid are dealt to the candidates in proportion to their number of chunks (a candidate in
    // two long contigs has many times the contacts of one in two short ones: equal shares would wait for the
    // largest); inside a candidate the items = (chunk, group of gs uniq slots) are strided over its warps.
    const int n_cands = sc->n_cands;
    int k = -1, b_first = 0, n_blocks_k = 0, tiles_k = 0;
    flat_block_range(sc, n_cands, (int)gridDim.x, -1, (int)blockIdx.x, &b_first, &n_blocks_k, &tiles_k, &k);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = lane; i < IG_N_OPS; i += 32) red[w][i] = 0.0;
    __syncwarp();
    if (k >= 0) {
    const int nw = n_blocks_k * IG_WARPS_PER_BLOCK;
    int gs = IG_N_OPS;
    while (gs > 3 && tiles_k * (IG_N_OPS / gs) < items_per_warp * nw) gs >>= 1;
    const int ng = IG_N_OPS / gs;
    const int n_items = tiles_k * ng;
    const int wg = ((int)blockIdx.x - b_first) * IG_WARPS_PER_BLOCK + w;
    for (int u = 0; u < IG_N_OPS; u++) acc_s[u * IG_THREADS + threadIdx.x] = 0.0;
    const Params p = sc->p;
    const double l10v = sc->log10_vinter;
    const double inter_const = (double)p.v_inter * LOG10E_F;
    double* my_acc = acc_s + threadIdx.x;
    QEnt* myq = queue[w];
    int qn = 0;
    unsigned touched = 0;
    const int n_uniq = desc_g[k].n_uniq;
    const int* cnt = flat_cnt + (size_t)k * chunk_stride;
    const FlatRec* my_flat = flat + (size_t)k * flat_stride;
    const RowMut* tab = table + (size_t)k * IG_N_OPS * ns;
    const int* tlen = table_len + (size_t)k * IG_N_OPS * ns;
    const IgMotion* g_mot = clstab[k].mot;
    TLP(12);  // prologue
    for (int it = wg; it < n_items; it += nw) {
        const int tile = it / ng, grp = it - tile * ng;
        const int u0 = grp * gs;
        if (u0 >= n_uniq) continue;
        const unsigned gmask = ((1u << gs) - 1u) << u0;
        unsigned m = 0;
        Ctc x;
        x.pos = 0; x.start_bp = 0; x.len_ori = 0; x.watson = 0.f; x.crick = 0.f; x.val = 0; x.cur_s = 0.f; x.cur_dp = 0;
        x.rjc = 0; x.t_cur = 0.0; x.flags = 0;
        int ri = 0;
        float row_s_tot = 0.f;
        if (lane < __ldg(&cnt[tile])) {
            const FlatRec r = my_flat[((size_t)tile << 5) + lane];
            x.pos = r.pos; x.start_bp = r.start_bp; x.len_ori = r.len_ori; x.watson = r.watson; x.crick = r.crick; x.val = r.val;
            x.cur_s = r.cur_s; x.cur_dp = r.cur_dp; x.rjc = r.rjc; x.t_cur = r.t_cur; x.flags = r.flags;
            ri = r.ri;
            m = r.m & gmask;
            row_s_tot = (r.flags & 4) ? 1.0f : 0.0f;   // eval_pair only asks whether the row's contig is circular
        }
        const unsigned um = __reduce_or_sync(0xffffffffu, m);
        if (!um) continue;
        unsigned chg = 0;
        RowMut a_nxt = tab[(size_t)(__ffs(um) - 1) * ns + ri];   // row-end entry, fetched one mutation ahead
#pragma unroll 1
        for (unsigned uw = um; uw; uw &= uw - 1) {
            const int u = __ffs(uw) - 1;
            const RowMut a = a_nxt;
            const unsigned rest = uw & (uw - 1);
            if (rest) a_nxt = tab[(size_t)(__ffs(rest) - 1) * ns + ri];
            float s_m = 0.f; int dp_m = 0; bool push = false;
            if ((m >> u) & 1u) {
                double add;
                if (eval_pair(x, u, a, g_mot, row_s_tot, p, l10v, inter_const, mbar, exz_tab, tab, tlen, ns, s_m, dp_m, push, add)) {
                    chg |= 1u << u;
                    my_acc[u * IG_THREADS] += add;
                }
            }
            queue_push(push, s_m, dp_m, 1u << u, x.val, myq, qn, my_acc, p, l10v, exz_tab);
        }
        queue_push((x.flags & 2) && chg, x.cur_s, x.cur_dp, chg | IG_QSUB, x.val, myq, qn, my_acc, p, l10v, exz_tab);
        touched |= __reduce_or_sync(0xffffffffu, chg);
    }
    TLP(13);  // items
    if (qn > 0) { eval_queue(myq, qn, my_acc, p, l10v, exz_tab); }
    __syncwarp();
    for (unsigned tw = touched; tw; tw &= tw - 1) {
        const int us = __ffs(tw) - 1;
        const double v = warp_sum(my_acc[us * IG_THREADS]);
        if (lane == 0) red[w][us] = v;
    }
    TLP(14);  // final flush + reductions
    }
    __syncthreads();
    if (k >= 0 && threadIdx.x < 25) {   // k_finalize reads candidate k's partials from its own block range only
        double v = 0.0;
        if (threadIdx.x < IG_N_OPS) for (int ww = 0; ww < IG_WARPS_PER_BLOCK; ww++) v += red[ww][threadIdx.x];
        part_nz[PART_IDX(k, 25, threadIdx.x, gridDim.x, blockIdx.x)] = v;
    }
}

__device__ void select_step(DevScalars* sc, const IgDescriptor* __restrict__ desc_g);

// K9: per-candidate finalisation: fixed-order parallel reduction of the block partials, the
//     reference's last-block quirk (KA:4362), zero terms (eval_all_likelihood_on_zero_2nd
//     KA:4005-4027) and score assembly (eval_all_scores KA:4029-4046).  One block per candidate.
__global__ void __launch_bounds__(1024)
k_finalize(const long long* __restrict__ row_ptr, const int2* __restrict__ cv, const CoordRec* __restrict__ coord,
           DevScalars* sc, const IgDescriptor* __restrict__ desc_g, const int* __restrict__ rows, const int* __restrict__ rowidx,
           int ns, const int* __restrict__ row_cnt, const RowMut* __restrict__ table, const int* __restrict__ table_len,
           float mbar, const float* __restrict__ exz_tab, const double* __restrict__ part_nz, const int* __restrict__ part_c,
           int n_part, const double* __restrict__ part_z, const int* __restrict__ part_i, int n_part_z, double n_pix,
           int compat_last_block, int* __restrict__ n_uniq_out, int* __restrict__ n_sub_out, int do_select, int n_part_c, int flat) {
    TL(7);
    TLP_DECL();
    const int k = blockIdx.x;
    if (k >= sc->n_cands) return;
    __shared__ double s_nz[25], s_z[25], s_corr[IG_N_OPS];
    __shared__ int s_i[25], s_c[2];
    __shared__ double t_val[IG_N_OPS][64];
    __shared__ int2 t_cv[64];
    __shared__ int t_ri[64];
    __shared__ int t_cnt, t_need, tr_n;
    __shared__ int tr_ri[64], tr_skip[64], tr_off[64];
    __shared__ int t_wsum[32];
    const IgDescriptor& d = desc_g[k];
    const Params p = sc->p;
    const double l10v = sc->log10_vinter;
    const CandInfo ci_k = sc->ci[k];
    const int n_uniq = d.n_uniq;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    // 25 + 25 + 25 + 2 slots, one warp per slot, lanes stride the partial blocks, fixed shuffle tree
    // 32 warps, one slot each per round; every lane first issues all of its loads (independent, in
    // flight together), then adds them in index order; fixed shuffle tree => deterministic
    int nz_first = 0, nz_count = n_part;
    if (flat) { int t_, c_; flat_block_range(sc, sc->n_cands, n_part, k, 0, &nz_first, &nz_count, &t_, &c_); }   // k_eval_flat's blocks for k
    for (int slot = w; slot < 77; slot += nwarp) {
        if (slot < 50) {
            const double* src = slot < 25 ? &part_nz[PART_IDX(k, 25, slot, n_part, nz_first)] : &part_z[PART_IDX(k, 25, slot - 25, n_part_z, 0)];
            const int n = slot < 25 ? nz_count : n_part_z;
            double v = 0.0;
            for (int i0 = 0; i0 < n; i0 += 32 * 16) {
                double x[16];
#pragma unroll
                for (int j = 0; j < 16; j++) { const int i = i0 + j * 32 + lane; x[j] = i < n ? src[i] : 0.0; }
#pragma unroll
                for (int j = 0; j < 16; j++) v += x[j];
            }
            v = warp_sum(v);
            if (lane == 0) { if (slot < 25) s_nz[slot] = v; else s_z[slot - 25] = v; }
        } else {
            // (the selection counters may come from another kernel than the likelihood partials: own block count)
            const int* src = slot < 75 ? &part_i[PART_IDX(k, 25, slot - 50, n_part_z, 0)] : &part_c[PART_IDX(k, 2, slot - 75, n_part_c, 0)];
            const int n = slot < 75 ? n_part_z : n_part_c;
            int v = 0;
            for (int i0 = 0; i0 < n; i0 += 32 * 16) {
                int x[16];
#pragma unroll
                for (int j = 0; j < 16; j++) { const int i = i0 + j * 32 + lane; x[j] = i < n ? src[i] : 0; }
#pragma unroll
                for (int j = 0; j < 16; j++) v += x[j];
            }
            v = __reduce_add_sync(0xffffffffu, v);
            if (lane == 0) { if (slot < 75) s_i[slot - 50] = v; else s_c[slot - 75] = v; }
        }
    }
    if (threadIdx.x < IG_N_OPS) s_corr[threadIdx.x] = 0.0;
    __syncthreads();
    TLP(5);
    const int n_sub = s_c[0];
    const int t = n_sub % 64;
    const RowMut* tab = table + (size_t)k * IG_N_OPS * ns;
    const int* tlen = table_len + (size_t)k * IG_N_OPS * ns;
    // ---- last-block quirk: uniq slots u >= t lose the final (n_sub % 64) contacts of the row-sorted slice
    if (compat_last_block && t > 0 && t < n_uniq) {
        // ordered (hence deterministic) collection of the last t selected contacts of the row-sorted slice:
        // rows from the last one backwards, the whole block scans a row's contacts with a block-wide
        // exclusive scan of the selection flags, keeping the row's last `take` selected contacts in order
        // 1. the tail rows, found in parallel: windows of blockDim rows from the end of the affected-row list,
        //    block-wide scan of their selected-contact counts (thread order = descending row)
        const int* rc_k = row_cnt + (size_t)k * ns;
        if (threadIdx.x == 0) { tr_n = 0; t_need = 0; }
        __syncthreads();
        for (int hi = ci_k.n_rows; hi > 0; hi -= (int)blockDim.x) {
            const int carry = t_need;   // selected contacts in the rows behind this window
            if (carry >= t) break;
            const int ri = hi - 1 - (int)threadIdx.x;
            const int c = ri >= 0 ? rc_k[ri] : 0;
            int x = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) t_wsum[w] = x;
            __syncthreads();
            int before = 0, total = 0;
            for (int ww = 0; ww < nwarp; ww++) { const int v = t_wsum[ww]; if (ww < w) before += v; total += v; }
            const int excl = carry + before + x - c;   // tail contacts in later rows
            if (c > 0 && excl < t) {
                const int take = min(c, t - excl);
                const int slot = atomicAdd(&tr_n, 1);  // < 64 rows: each holds at least one tail contact
                tr_ri[slot] = ri; tr_skip[slot] = c - take; tr_off[slot] = t - (excl + take);
            }
            __syncthreads();
            if (threadIdx.x == 0) t_need = carry + total;
            __syncthreads();
        }
        // 2. one warp per tail row: its last `take` selected contacts, in column order, to their place in the tail
        for (int j = w; j < tr_n; j += nwarp) {
            const int ri = tr_ri[j], skip = tr_skip[j], off = tr_off[j];
            const int r = rows[(size_t)k * ns + ri];
            const CoordRec ci = coord[r];
            const long long b0 = row_ptr[r], e0 = row_ptr[r + 1];
            int running = 0;
            for (long long q0 = b0; q0 < e0; q0 += 32) {
                const long long q = q0 + lane;
                int2 c = make_int2(0, 0);
                bool sel = false;
                if (q < e0) {
                    c = cv[q];
                    const CoordRec cj = coord[c.x];
                    sel = (cj.id_c == ci_k.id_a || cj.id_c == ci_k.id_b) && contact_selected(ci, cj, c.y, ci_k);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, sel);
                const int idx = running + __popc(bal & ((1u << lane) - 1));
                if (sel && idx >= skip) { const int slot = off + (idx - skip); t_cv[slot] = c; t_ri[slot] = ri; }
                running += __popc(bal);
            }
        }
        if (threadIdx.x == 0) t_cnt = t;
        __syncthreads();
        const int n_items = t_cnt * (n_uniq - t);
        for (int idx = threadIdx.x; idx < n_items; idx += blockDim.x) {
            const int e = idx % t_cnt, u = t + idx / t_cnt;
            const int2 c = t_cv[e];
            const RowMut a = tab[(size_t)u * ns + t_ri[e]];
            const int rj = rowidx[(size_t)k * ns + c.x] & ((1 << IG_CLS_SHIFT) - 1);
            const RowMut bm = tab[(size_t)u * ns + rj];
            CoordRec cim, cjm;
            cim.dist = a.dist; cim.id_c = a.id_c; cim.pos = a.pos; cim.s_tot = a.s_tot;
            cjm.dist = bm.dist; cjm.id_c = bm.id_c; cjm.pos = bm.pos; cjm.s_tot = bm.s_tot;
            const double ob = (double)c.y;
            t_val[u][e] = contact_term(cim, cjm, tlen[(size_t)u * ns + rj], ob, ob_const(ob), p, l10v, mbar, exz_tab);
        }
        __syncthreads();
        if (threadIdx.x >= t && threadIdx.x < n_uniq) {
            double ssum = 0.0;
            for (int e = 0; e < t_cnt; e++) ssum += t_val[threadIdx.x][e];
            s_corr[threadIdx.x] = ssum;
        }
        __syncthreads();
    }
    TLP(6);
    // ---- scores
    if (threadIdx.x < IG_N_OPS) sc->scores[k * IG_N_OPS + threadIdx.x] = 0.0;
    __syncthreads();
    if (threadIdx.x < n_uniq) {
        const int u = threadIdx.x;
        const double log_e = (double)LOG10E_F;
        const int m = d.uniq[u];
        // Z[m] over ALL sub-fragments = Z_cur(all) - Z_cur(affected rows) + Z_m(affected rows)
        const double z = sc->z_cur - s_z[24] + s_z[u];
        const int n_intra = sc->nintra_cur - s_i[24] + s_i[u];  // int32 wrap-consistent
        const double val_inter = -1.0 * log_e * (n_pix - __int2double_rn(n_intra)) * p.v_inter;
        const double lz = z * log_e + val_inter;
        const double lnz = s_nz[u] - ((u >= t && compat_last_block && t > 0) ? s_corr[u] : 0.0);
        sc->scores[k * IG_N_OPS + m] = lnz + lz + sc->lnz_full - s_nz[24];
        sc->lnz_new[k * IG_N_OPS + m] = sc->lnz_full - s_nz[24] + s_nz[u];  // without the last-block quirk
        sc->z_new[k * IG_N_OPS + m] = z;
        sc->nintra_new[k * IG_N_OPS + m] = n_intra;
    }
    if (threadIdx.x == 0) {
        sc->lsub_cur[k] = s_nz[24];
        sc->ci[k].n_sub = n_sub;
        n_uniq_out[k] = n_uniq;
        n_sub_out[k] = n_sub;
        atomicAdd(&sc->st_contacts, (unsigned long long)s_c[1]);
        atomicAdd(&sc->st_rows, (unsigned long long)ci_k.n_rows);
        atomicAdd(&sc->st_frags, (unsigned long long)(d.A.l_cont + (ci_k.same ? 0 : d.B.l_cont)));
        atomicAdd(&sc->st_selected, (unsigned long long)n_sub);
        atomicAdd(&sc->st_proposals, (unsigned long long)n_uniq);
    }
    TLP(7);
    if (!do_select) return;
    // move selection by the LAST candidate block to finish (saves a dependent launch)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) t_cnt = (atomicAdd(&sc->ticket_fin, 1u) == (unsigned)sc->n_cands - 1) ? 1 : 0;
    __syncthreads();
    if (t_cnt && threadIdx.x < 32) {
        __threadfence();
        select_step(sc, desc_g);
    }
}

// K10: move selection (CL:1435-1446): scores==0 -> -inf; first index of the maximum.
__global__ void k_select(DevScalars* sc) {
    __shared__ double sm[IG_MAX_CANDS * IG_N_OPS];
    const int n = sc->n_cands * IG_N_OPS;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = sc->scores[i];
    __syncthreads();
    if (threadIdx.x != 0) return;
    int best = -1;
    double bv = 0.0;
    for (int i = 0; i < n; i++) {
        const double v = sm[i];
        if (v == 0.0) continue;
        if (best < 0 || v > bv) { best = i; bv = v; }
    }
    if (best < 0) best = 0;  // np.argmax of an all-zero filtered vector
    sc->win_cand = best / IG_N_OPS;
    sc->win_op = best % IG_N_OPS;
    sc->likelihood = sm[best];
    sc->n_heads = 0; sc->sum_l_cont = 0; sc->dist_half = 0;  // accumulators of k_post
}
// step path: selection + the bookkeeping of k_post_scalars in one launch (k_apply reads the label base
// from the descriptor, not from sc->max_label, so bumping it here cannot race)
__device__ void select_step(DevScalars* sc, const IgDescriptor* __restrict__ desc_g) {
    const int n = sc->n_cands * IG_N_OPS;
    const int lane = threadIdx.x & 31;  // executed by one full warp
    // first index of the maximum among the scored (non-zero) proposals (CL:1435-1446)
    int best = -1;
    double bv = 0.0;
    for (int i = lane; i < n; i += 32) {
        const double v = __ldcg(&sc->scores[i]);
        if (v == 0.0) continue;
        if (best < 0 || v > bv) { best = i; bv = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int ob = __shfl_down_sync(0xffffffffu, best, o);
        const double ov = __shfl_down_sync(0xffffffffu, bv, o);
        if (ob >= 0 && (best < 0 || ov > bv || (ov == bv && ob < best))) { best = ob; bv = ov; }
    }
    best = __shfl_sync(0xffffffffu, best, 0);
    if (best < 0) best = 0;
    const int kc = best / IG_N_OPS, op = best % IG_N_OPS;
    const unsigned hit = __ballot_sync(0xffffffffu, lane < desc_g[kc].n_uniq && desc_g[kc].uniq[lane] == op);
    if (lane < 12 && op >= 12) sc->valid[lane] = desc_g[kc].valid[lane];
    if (lane != 0) return;
    sc->win_cand = kc; sc->win_op = op; sc->likelihood = __ldcg(&sc->scores[best]);
    sc->n_heads = 0; sc->sum_l_cont = 0; sc->dist_half = 0;
    sc->max_label += 2;
    sc->prev_k = kc; sc->prev_u = hit ? (__ffs(hit) - 1) : 0;
    sc->prev_windowed = (sc->ci[kc].same && sc->ci[kc].is_circ == 0) ? 1 : 0;
    sc->prev_id_a = sc->ci[kc].id_a; sc->prev_n_rows = sc->ci[kc].n_rows;
    sc->lnz_next = __ldcg(&sc->lnz_new[best]); sc->z_next = __ldcg(&sc->z_new[best]); sc->nintra_next = __ldcg(&sc->nintra_new[best]);
    sc->ticket_out = 0;
}
__global__ void k_select_step(DevScalars* sc, const IgDescriptor* __restrict__ desc_g) { select_step(sc, desc_g); }

// Same-linear-contig moves are scored on a WINDOWED slice (slice_sp_mat, KA:565-586): contacts of the
// contig outside the windows keep their distance mathematically, but the reference's next full
// recomputation (CL:1409) sees their float32 coordinates re-rounded.  To keep lnz_full identical to
// that recomputation without rescanning every contact, add exactly those contacts' term changes.
__global__ void __launch_bounds__(IG_THREADS)
k_lnz_outside(const long long* __restrict__ row_ptr, const int2* __restrict__ cv, const CoordRec* __restrict__ coord,
              const int* __restrict__ clen, DevScalars* sc, const int* __restrict__ rows, const int* __restrict__ rowidx, int ns,
              const RowMut* __restrict__ table, const int* __restrict__ table_len, float mbar, const float* __restrict__ exz_tab,
              double* __restrict__ part) {
    TL(8);
    if (!sc->prev_windowed) return;
    __shared__ double sm[32];
    __shared__ int is_last;
    const int k = sc->prev_k, u = sc->prev_u;
    const Params p = sc->p;
    const double l10v = sc->log10_vinter;
    const CandInfo ci_k = sc->ci[k];
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int* my_rows = rows + (size_t)k * ns;
    const int* my_idx = rowidx + (size_t)k * ns;
    const RowMut* tab = table + ((size_t)k * IG_N_OPS + u) * ns;
    const int* tlen = table_len + ((size_t)k * IG_N_OPS + u) * ns;
    double acc = 0.0;
    for (int ri = wg; ri < ci_k.n_rows; ri += nw) {
        const int r = my_rows[ri];
        const CoordRec ci = coord[r];
        const RowMut a = tab[ri];
        for (long long q = row_ptr[r] + lane; q < row_ptr[r + 1]; q += 32) {
            const int2 c = __ldg(&cv[q]);
            const CoordRec cj = coord[c.x];
            if (cj.id_c != ci_k.id_a) continue;               // other contigs: inter-contig term, unchanged
            if (contact_selected(ci, cj, c.y, ci_k)) continue;  // already inside lnz_new
            const int rj = my_idx[c.x] & ((1 << IG_CLS_SHIFT) - 1);
            const RowMut bm = tab[rj];
            CoordRec cim, cjm;
            cim.dist = a.dist; cim.id_c = a.id_c; cim.pos = a.pos; cim.s_tot = a.s_tot;
            cjm.dist = bm.dist; cjm.id_c = bm.id_c; cjm.pos = bm.pos; cjm.s_tot = bm.s_tot;
            const double ob = (double)c.y, obc = ob_const(ob);
            // the full-likelihood kernel takes the circular length from the ROW (KA:4428)
            const double t_old = contact_term(ci, cj, clen[r], ob, obc, p, l10v, mbar, exz_tab);
            const double t_new = contact_term(cim, cjm, tlen[ri], ob, obc, p, l10v, mbar, exz_tab);
            acc += t_new - t_old;
        }
    }
    const double tot = block_sum(acc, sm);
    if (threadIdx.x == 0) {
        part[blockIdx.x] = tot;
        __threadfence();
        is_last = (atomicAdd(&sc->ticket_out, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double v = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v += ((volatile double*)part)[i];
    const double all = block_sum(v, sm);
    if (threadIdx.x == 0) sc->lnz_next += all;
}

// start of the next step in incremental mode: the coordinates of the rows touched by the last applied
// move are taken from the mutation table (bit-identical to uni_fill_vect_dist on the new scaffold),
// everything else is unchanged; the scalar likelihood pieces were prepared by the previous step.
__global__ void __launch_bounds__(256)
k_commit_coords(CoordRec* __restrict__ coord, int* __restrict__ clen, DevScalars* sc, const int* __restrict__ rows, int ns,
                const RowMut* __restrict__ table, const int* __restrict__ table_len, const FragRec* __restrict__ live,
                const SubRec* __restrict__ sub, SubX* __restrict__ subx) {
    TL(11);
    const int k = sc->prev_k, u = sc->prev_u, n = sc->prev_n_rows;
    const int* my_rows = rows + (size_t)k * ns;
    const RowMut* tab = table + ((size_t)k * IG_N_OPS + u) * ns;
    const int* tlen = table_len + ((size_t)k * IG_N_OPS + u) * ns;
    for (int ri = blockIdx.x * blockDim.x + threadIdx.x; ri < n; ri += gridDim.x * blockDim.x) {
        const int r = my_rows[ri];
        const RowMut m = tab[ri];
        CoordRec c; c.dist = m.dist; c.id_c = m.id_c; c.pos = m.pos; c.s_tot = m.s_tot;
        coord[r] = c; clen[r] = tlen[ri];
        const SubRec sr = sub[r];
        const Frag f = live[sr.parent].f;   // the scaffold after the applied move
        SubX x; x.start_bp = f.start_bp; x.len_ori = f.len_bp * f.ori; x.watson = sr.watson; x.crick = sr.crick;
        subx[r] = x;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        sc->lnz_full = sc->lnz_next; sc->z_cur = sc->z_next; sc->nintra_cur = sc->nintra_next;
    }
}

// K11: apply the winning move to every fragment (test_copy_struct + copy_struct, CL:2094-2151)
__global__ void __launch_bounds__(256)
k_apply(FragRec* __restrict__ live, int nf, DevScalars* sc, const IgDescriptor* __restrict__ desc_g, int forced_cand, int forced_op) {
    TL(9);
    __shared__ IgDescriptor d;
    const int kc = forced_cand >= 0 ? forced_cand : sc->win_cand;
    const int op = forced_op >= 0 ? forced_op : sc->win_op;
    {
        const int* src = reinterpret_cast<const int*>(desc_g + kc);
        int* dst = reinterpret_cast<int*>(&d);
        for (int i = threadIdx.x; i < (int)(sizeof(IgDescriptor) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    // every thread reads only its own fragment + the descriptor's pivots (loaded before any write): in place is safe
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nf) {
        const Frag f = live[i].f;
        Frag o;
        if (op >= 8 && op < 12) {  // paste may leave a fragment unwritten (Q4): keep + count
            const int ua = (op - 8) >> 1, ub = (op - 8) & 1;
            Frag t1 = ig_split(f, i, d.A, ua, d.max_id);
            Frag t2 = ig_split(t1, i, d.T1B[ua], ub, d.max_id1[ua]);
            int written;
            o = ig_paste(t2, i, d.T2A[ua][ub], d.a, d.T2B[ua][ub], d.b, &written);
            if (!written) atomicAdd(&sc->q4_hits, 1);
        } else {
            o = ig_eval_op(d, op, f, i);
        }
        live[i].f = o;
    }
}
// bookkeeping that must not race with k_apply's reads of sc->max_label (through the descriptor it does not: the
// descriptor carries max_id) -- label counter and list_valid_insert (CL:2125-2126 re-runs get_bounds for ops >= 12)
__global__ void k_post_scalars(DevScalars* sc, const IgDescriptor* __restrict__ desc_g, int forced_cand, int forced_op) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int kc = forced_cand >= 0 ? forced_cand : sc->win_cand;
    const int op = forced_op >= 0 ? forced_op : sc->win_op;
    if (op >= 12) for (int i = 0; i < 12; i++) sc->valid[i] = desc_g[kc].valid[i];
    sc->max_label += 2;
    if (forced_cand >= 0) { sc->n_heads = 0; sc->sum_l_cont = 0; sc->dist_half = 0; }
}
__global__ void __launch_bounds__(256)
k_post(const FragRec* __restrict__ live, int nf, const int* __restrict__ init_prev, const int* __restrict__ init_next,
       const int* __restrict__ orientable, DevScalars* sc, CycleOut* __restrict__ cyc_out, const int* __restrict__ d_nuniq,
       const int* __restrict__ d_nsub) {
    TL(10);
    __shared__ int s_heads;
    __shared__ long long s_len, s_half;
    if (threadIdx.x == 0) { s_heads = 0; s_len = 0; s_half = 0; }
    __syncthreads();
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    int heads = 0; long long len = 0; int half = 0;  // half = decrement of d in units of 1/2
    if (f < nf) {
        const Frag g = live[f].f;
        if (g.pos == 0) { heads = 1; len = g.l_cont; }
        // dist_inter_genome, CL:672-715 (init_ori == +1, blacklist empty)
        const int p0 = init_prev[f], n0 = init_next[f];
        int p1 = g.prev, n1 = g.next;
        int swap = 1;
        if ((p1 == p0 && n1 == n0) || (p1 == n0 && n1 == p0)) half += 2;
        if (orientable[f]) {
            if (1 != g.ori) { int tmp = p1; p1 = n1; n1 = tmp; swap = -1; }
            if (p0 == p1) {
                if (p0 == -1 || !orientable[p1]) half += 2;
                else { half += 1; if (1 == swap * live[p1].f.ori) half += 1; }
            }
            if (n0 == n1) {
                if (n0 == -1 || !orientable[n1]) half += 2;
                else { half += 1; if (1 == swap * live[n1].f.ori) half += 1; }
            }
        } else {
            if (p1 == p0 || p1 == n0) half += 2;
            if (n1 == n0 || n1 == p0) half += 2;
        }
    }
    // warp-level reductions first: one shared atomic per warp instead of one per thread
    const int w_heads = __reduce_add_sync(0xffffffffu, heads);
    const int w_half = __reduce_add_sync(0xffffffffu, half);
    long long w_len = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w_len += __shfl_down_sync(0xffffffffu, w_len, o);
    if ((threadIdx.x & 31) == 0) {
        if (w_heads) { atomicAdd(&s_heads, w_heads); atomicAdd((unsigned long long*)&s_len, (unsigned long long)w_len); }
        if (w_half) atomicAdd((unsigned long long*)&s_half, (unsigned long long)w_half);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_heads) atomicAdd(&sc->n_heads, s_heads);
        if (s_len) atomicAdd((unsigned long long*)&sc->sum_l_cont, (unsigned long long)s_len);
        if (s_half) atomicAdd((unsigned long long*)&sc->dist_half, (unsigned long long)s_half);
        if (cyc_out) {  // cycle mode: the last block to finish publishes this step's record and advances the plan
            __threadfence();
            if (atomicAdd(&sc->ticket_post, 1u) == gridDim.x - 1) {
                __threadfence();
                CycleOut o;
                o.likelihood = sc->likelihood; o.lnz_full = sc->lnz_full;
                o.dist_half = *(volatile long long*)&sc->dist_half; o.sum_l_cont = *(volatile long long*)&sc->sum_l_cont;
                o.n_heads = *(volatile int*)&sc->n_heads; o.win_cand = sc->win_cand; o.win_op = sc->win_op; o.q4_hits = sc->q4_hits;
                for (int i = 0; i < IG_MAX_CANDS; i++) { o.n_uniq[i] = d_nuniq[i]; o.n_sub[i] = d_nsub[i]; o.pad[i] = 0; }
                cyc_out[sc->step_idx] = o;
                sc->step_idx += 1;
#ifdef IG_TIMELINE
                g_tl_step += 1;
#endif
            }
        }
    }
}
__global__ void k_explode(FragRec* live, int nf, const int* __restrict__ perm) {  // KA:409-426
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf) return;
    Frag f = live[i].f;
    f.pos = 0; f.start_bp = 0; f.sub_pos = 0; f.id_c = perm[i]; f.prev = -1; f.next = -1;
    f.l_cont = 1; f.l_cont_bp = f.len_bp; f.sub_l_cont = f.sub_len;
    live[i].f = f;
}
// histogram for the initial p(s) fit (CL:2253-2293) on the INITIAL scaffold; integer-exact sums
__global__ void __launch_bounds__(IG_THREADS)
k_histogram(const long long* __restrict__ row_ptr, const int2* __restrict__ cv, const int* __restrict__ sym_diag,
            const FragRec* __restrict__ init, const SubRec* __restrict__ sub, int n_rows, double bin_kb, double max_kb,
            int n_bins, unsigned long long* __restrict__ hist, unsigned long long* __restrict__ rows_used) {
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = wg; r < n_rows; r += nw) {
        const SubRec si = sub[r];
        const Frag fi = init[si.parent].f;
        const bool used = bin_kb < (double)fi.l_cont_bp / 1000.0;
        if (!used) continue;
        const double s_i = (double)fi.start_bp / 1000.0 + (double)si.watson;
        if (lane == 0) {
            atomicAdd(rows_used, 1ULL);
            if (sym_diag && sym_diag[r] != 0 && 0.0 < max_kb) atomicAdd(&hist[0], (unsigned long long)sym_diag[r]);
        }
        for (long long q = row_ptr[r] + lane; q < row_ptr[r + 1]; q += 32) {
            const int2 c = cv[q];
            const SubRec sj = sub[c.x];
            const Frag fj = init[sj.parent].f;
            if (fj.id_c != fi.id_c) continue;
            const double s_j = (double)fj.start_bp / 1000.0 + (double)sj.watson;
            const double dd = fabs(s_i - s_j);
            if (!(dd < max_kb)) continue;
            const int b = (int)(dd / bin_kb);
            if (b < 0 || b >= n_bins) continue;
            const int mult = 1 + (c.x < n_rows ? 1 : 0);  // symmetric matrix: row r and row c.x both see it
            atomicAdd(&hist[b], (unsigned long long)((long long)c.y * mult));
        }
    }
}

// N1: K x K thumbnail of the contact map in the CURRENT scaffold order, binned on the
// device (the reference densifies NS x NS on the host, CL:2598-2599).  Integer counts => exact.
__global__ void __launch_bounds__(IG_THREADS)
k_thumbnail(const long long* __restrict__ row_ptr, const int2* __restrict__ cv, const int* __restrict__ sub_rank, int ns, int K,
            unsigned int* __restrict__ img) {
    const int lane = threadIdx.x & 31;
    const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = wg; r < ns; r += nw) {
        const int pi = (int)(((long long)sub_rank[r] * K) / ns);
        for (long long q = row_ptr[r] + lane; q < row_ptr[r + 1]; q += 32) {
            const int2 c = cv[q];
            const int pj = (int)(((long long)sub_rank[c.x] * K) / ns);
            atomicAdd(&img[(size_t)pi * K + pj], (unsigned int)c.y);
            if (pi != pj) atomicAdd(&img[(size_t)pj * K + pi], (unsigned int)c.y);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Production RNG mode: the neighbour draws of a whole cycle on the device.
// Philox4x32-10 keyed by the seed, counter = (step, cycle, draw index, attempt); one thread per step.
// Distribution = return_neighbours (CL:3103-3141): min(delta, #non-zero weights) fragments drawn without
// replacement with probability proportional to the level's contact counts (successive draws, a drawn
// fragment is rejected when drawn again), or `delta` distinct uniform fragments when A has no neighbour;
// then sorted (CL:1404) and A itself dropped.  The stream differs from NumPy's by design
// (the tests hold a NumPy restatement of this kernel).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ double philox_uniform(unsigned step, unsigned cycle, unsigned draw, unsigned attempt, uint2 key) {
    const uint4 r = philox4x32_10(make_uint4(step, cycle, draw, attempt), key);
    const unsigned long long bits = ((unsigned long long)r.x << 32) | r.y;
    return (double)(bits >> 11) * (1.0 / 9007199254740992.0);  // 53 bits -> [0, 1)
}
__global__ void k_draw_plan(int* __restrict__ plan, const int* __restrict__ frags, int n_steps, int delta, int nf,
                            const long long* __restrict__ nb_ptr, const int* __restrict__ nb_idx, const double* __restrict__ nb_cdf,
                            const int* __restrict__ nb_nnz, unsigned seed_lo, unsigned seed_hi, unsigned cycle) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_steps) return;
    const uint2 key = make_uint2(seed_lo, seed_hi);
    const int a = frags[t];
    const long long b = nb_ptr[a], e = nb_ptr[a + 1];
    const int len = (int)(e - b);
    int got[IG_MAX_CANDS];
    int n = 0;
    if (len > 0) {
        const int n_max = min(delta, nb_nnz[a]);
        const double total = nb_cdf[e - 1];
        for (int i = 0; i < n_max; i++) {
            int pick = -1;
            for (unsigned att = 0; att < 256u && pick < 0; att++) {
                const double u = philox_uniform((unsigned)t, cycle, (unsigned)i, att, key) * total;
                int lo = 0, hi = len - 1;  // first j with cdf[j] > u
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (nb_cdf[b + mid] > u) hi = mid; else lo = mid + 1; }
                const int c = nb_idx[b + lo];
                bool dup = false;
                for (int j = 0; j < n; j++) dup |= (got[j] == c);
                if (!dup) pick = c;
            }
            if (pick < 0) {  // a weight so dominant that 256 redraws all hit it: take the first free non-zero entry
                for (int j = 0; j < len && pick < 0; j++) {
                    const double wj = nb_cdf[b + j] - (j ? nb_cdf[b + j - 1] : 0.0);
                    const int c = nb_idx[b + j];
                    bool dup = false;
                    for (int q = 0; q < n; q++) dup |= (got[q] == c);
                    if (wj > 0.0 && !dup) pick = c;
                }
            }
            if (pick >= 0) got[n++] = pick;
        }
    } else {
        const int n_max = min(delta, nf - 1);
        for (int i = 0; i < n_max; i++) {
            int pick = -1;
            for (unsigned att = 0; att < 256u && pick < 0; att++) {
                const int c = min(nf - 1, (int)(philox_uniform((unsigned)t, cycle, (unsigned)i, att, key) * (double)nf));
                bool dup = (c == a);
                for (int j = 0; j < n; j++) dup |= (got[j] == c);
                if (!dup) pick = c;
            }
            if (pick >= 0) got[n++] = pick;
        }
    }
    // sorted, without A itself
    for (int i = 1; i < n; i++) { const int v = got[i]; int j = i - 1; while (j >= 0 && got[j] > v) { got[j + 1] = got[j]; j--; } got[j + 1] = v; }
    int* p = plan + (size_t)t * (2 + IG_MAX_CANDS);
    int m = 0;
    for (int i = 0; i < n; i++) if (got[i] != a) p[2 + m++] = got[i];
    for (int i = m; i < IG_MAX_CANDS; i++) p[2 + i] = 0;
    p[0] = m; p[1] = a;
}

// ================================================================================================
// host side
struct ig_handle {
    ig_config cfg;
    int nf, ns;
    long long nnz;
    cudaStream_t stream, side, pf;
    cudaEvent_t ev_coords, ev_lnz, ev_fork, ev_sel, ev_out, ev_cuts, ev_cls;
    bool rows_small;  // affected-row list in one launch (small levels)
    bool flat;        // flat scoring path (k_pick + k_eval_flat) for small levels
    int grid_flat, flat_items; int* flat_cnt; FlatRec* flat_list; size_t flat_stride, chunk_stride;
    bool prefetch;  // the level's arrays fit the L2 comfortably: prefetch them at the start of a step
    FragRec *live, *init_live;
    SubRec* sub;
    CoordRec* coord;
    int* clen;
    long long* row_ptr;
    int2* cv;
    int* sym_diag;
    int *init_prev, *init_next, *orientable;
    DevScalars* sc;
    IgDescriptor* desc;
    float *exz, *exz_test;
    int n_chunks, *chunk_cnt, *rows, *row_cnt;
    int grid_score;
    double *part_nz, *part_z; int *part_i, *part_c;
    int grid_pre;
    RowMut* table; int *table_len, *rowidx;
    IgClassTab* clstab; int rigid; SubX* subx; RowInfo* rinfo;
    long long* nb_ptr; int* nb_idx; double* nb_cdf; int* nb_nnz; int* cyc_frags;
    double *part_full; int n_part_full;
    double *part_zc; int* part_nc; int n_part_zc;
    int *d_nuniq, *d_nsub, *d_perm;
    unsigned long long* d_hist;
    DevScalars* h_sc;  // pinned mirror
    int* h_small;      // pinned scratch (cands, nuniq, nsub)
    bool params_set, coords_fresh, coords_ever;
    bool incr_valid; int refresh_every; long long steps_since_full;
    double* part_out;
    int gs_div, sparse_div, grid_split;
    long long last_n_full;
    cudaGraphExec_t graph[4][IG_MAX_CANDS + 1];   // [full + 2 * cycle][candidates in the grid]
    int* cyc_in; CycleOut* cyc_out; int cyc_cap; bool graph_failed, capturing, use_graph; long long n_full;
    // measurement (CUDA events on the launching stream)
    cudaEvent_t ev[6];
    cudaEvent_t evk[16]; double ms_k[16];  // per-kernel event timing of the main stream (profiling mode)
    double ms_step, ms_score, ms_full;
    long long n_launches, n_steps;
    int profile;
    std::string err;
};

static thread_local std::string g_err;
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            char buf[512];                                                                         \
            snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            if (h) h->err = buf; else g_err = buf;                                                 \
            return -2;                                                                             \
        }                                                                                          \
    } while (0)

extern "C" const char* ig_last_error(ig_handle* h) { return h ? h->err.c_str() : g_err.c_str(); }
extern "C" int ig_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

template <class T> static int dev_alloc(ig_handle* h, T** p, size_t n) {
    CK(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    return 0;
}

static int upload_state(ig_handle* h, const int32_t* in13, FragRec* dst) {
    const int nf = h->nf;
    std::vector<FragRec> tmp(nf);
    for (int i = 0; i < nf; i++) {
        int* v = reinterpret_cast<int*>(&tmp[i].f);
        for (int k = 0; k < IG_N_FIELDS; k++) v[k] = in13[(size_t)k * nf + i];
        tmp[i].pad[0] = tmp[i].pad[1] = tmp[i].pad[2] = 0;
    }
    CK(cudaMemcpyAsync(dst, tmp.data(), sizeof(FragRec) * nf, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int ig_create(const ig_config* cfg, const ig_level_data* data, ig_handle** out) {
    ig_handle* h = nullptr;
    if (!cfg || !data || !out) { g_err = "ig_create: null argument"; return -1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        g_err = "ig_create: no CUDA device available (this library has no CPU path)";
        return -3;
    }
    if (cfg->device < 0 || cfg->device >= ndev) { g_err = "ig_create: bad device ordinal"; return -1; }
    if (cfg->n_frags <= 0 || cfg->n_sub_frags <= 0 || cfg->nnz < 0) { g_err = "ig_create: bad sizes"; return -1; }
    if (cfg->n_sub_frags >= (1 << IG_CLS_SHIFT)) { g_err = "ig_create: more than 2^24 sub-fragments"; return -1; }
    h = new ig_handle();
    h->cfg = *cfg; h->nf = cfg->n_frags; h->ns = cfg->n_sub_frags; h->nnz = cfg->nnz;
    h->rigid = cfg->rigid_pruning ? 1 : 0;
    h->params_set = false; h->coords_fresh = false; h->coords_ever = false;
    h->incr_valid = false; h->refresh_every = 4096; h->steps_since_full = 0;
    h->gs_div = 4;
    h->sparse_div = 4;
    h->grid_split = 0;
    if (const char* e = getenv("IG_GRID_SPLIT")) h->grid_split = atoi(e);   // 0: the full grid per candidate (experiments)
    if (const char* e = getenv("IG_SPARSE_DIV")) h->sparse_div = atoi(e) & 0xffff;
    if (const char* e = getenv("IG_FORCE_SPLIT")) {  // "gs,parts" (experiments)
        int g = 0, pp = 0;
        if (sscanf(e, "%d,%d", &g, &pp) == 2 && g > 0 && pp > 0 && IG_N_OPS % g == 0) h->sparse_div |= (g << 16) | (pp << 24);
    }
    if (const char* e = getenv("IG_GS_DIV")) h->gs_div = std::max(1, atoi(e));
    if (const char* e = getenv("IG_BLOCK_MODE")) if (atoi(e)) h->gs_div = -h->gs_div;
    memset(h->graph, 0, sizeof h->graph); h->cyc_in = nullptr; h->cyc_out = nullptr; h->cyc_frags = nullptr; h->cyc_cap = 0;
    h->nb_ptr = nullptr; h->nb_idx = nullptr; h->nb_cdf = nullptr; h->nb_nnz = nullptr; h->graph_failed = false; h->capturing = false; h->use_graph = true; h->n_full = 0;
    cudaError_t e0 = cudaSetDevice(cfg->device);
    if (e0 != cudaSuccess) { g_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e0); delete h; return -2; }
#define CKC(x) do { int r_ = (x); if (r_) { g_err = h->err; ig_destroy(h); return r_; } } while (0)
    auto body = [&]() -> int {
        CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->pf, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_cuts, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_cls, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_coords, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_lnz, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_sel, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
        for (int i = 0; i < 6; i++) CK(cudaEventCreate(&h->ev[i]));
        for (int i = 0; i < 16; i++) { CK(cudaEventCreate(&h->evk[i])); h->ms_k[i] = 0.0; }
        h->ms_step = h->ms_score = h->ms_full = 0.0; h->n_launches = 0; h->n_steps = 0; h->profile = 0;
        const int nf = h->nf, ns = h->ns;
        if (dev_alloc(h, &h->live, nf) || dev_alloc(h, &h->init_live, nf)) return -2;
        if (dev_alloc(h, &h->sub, ns) || dev_alloc(h, &h->coord, ns) || dev_alloc(h, &h->clen, ns)) return -2;
        if (dev_alloc(h, &h->row_ptr, (size_t)ns + 1) || dev_alloc(h, &h->cv, (size_t)h->nnz)) return -2;
        if (dev_alloc(h, &h->init_prev, nf) || dev_alloc(h, &h->init_next, nf) || dev_alloc(h, &h->orientable, nf)) return -2;
        if (dev_alloc(h, &h->sc, 1) || dev_alloc(h, &h->desc, IG_MAX_CANDS)) return -2;
        if (dev_alloc(h, &h->exz, (size_t)ns + 1) || dev_alloc(h, &h->exz_test, (size_t)ns + 1)) return -2;
        h->n_chunks = (ns + IG_ROW_CHUNK - 1) / IG_ROW_CHUNK;
        if (dev_alloc(h, &h->chunk_cnt, (size_t)IG_MAX_CANDS * h->n_chunks)) return -2;
        h->rows_small = h->n_chunks <= IG_ROWS_SMALL_CHUNKS;
        if (const char* e = getenv("IG_ROWS_SMALL")) h->rows_small = h->rows_small && atoi(e) != 0;   // tests: force the two-pass list
        h->flat = h->rows_small && h->nnz <= IG_FLAT_MAX_NNZ;
        if (const char* e = getenv("IG_FLAT")) h->flat = h->flat && atoi(e) != 0;
        if (getenv("IG_FORCE_SPLIT")) h->flat = false;   // experiments / tests of the row-per-warp kernel
        h->flat_cnt = nullptr; h->flat_list = nullptr; h->flat_stride = (size_t)h->nnz + 32 * (size_t)ns + 32;   // rows own whole 32-contact chunks
        h->chunk_stride = h->flat_stride / 32 + 1;
        if (h->flat) {
            if (dev_alloc(h, &h->flat_cnt, (size_t)IG_MAX_CANDS * h->chunk_stride) ||
                dev_alloc(h, &h->flat_list, (size_t)IG_MAX_CANDS * h->flat_stride)) return -2;
        }
        if (dev_alloc(h, &h->rows, (size_t)IG_MAX_CANDS * ns) || dev_alloc(h, &h->row_cnt, (size_t)IG_MAX_CANDS * ns)) return -2;
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, cfg->device));
        const int sms = prop.multiProcessorCount;
        h->grid_score = sms * IG_SCORE_CTAS_PER_SM;  // 24 resident warps per SM (80 registers per thread; 6 KB of accumulators per warp)
        CK(cudaFuncSetAttribute(k_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IG_SCORE_SMEM));
        CK(cudaFuncSetAttribute(k_eval_flat, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IG_SCORE_SMEM));
        h->grid_flat = (sms * IG_SCORE_CTAS_PER_SM) / 5;   // k_pick grid per candidate: a step's (usually 5) candidates fill the GPU once
        h->flat_items = 2;
        if (const char* e = getenv("IG_FLAT_ITEMS")) h->flat_items = std::max(1, atoi(e));   // a step's (usually 5) candidates fill the GPU once
        h->grid_pre = std::min(sms * 2, (ns + 31) / 32);   // 2 resident CTAs of 25 warps per SM; one tile of 32 rows per block at most
        {
            const char* e = getenv("IG_PREFETCH");
            const size_t bytes = sizeof(int2) * (size_t)h->nnz + 64 * (size_t)ns + 80 * (size_t)nf;
            h->prefetch = e ? (atoi(e) != 0 && bytes <= ((size_t)48 << 20)) : false;
        }
        if (dev_alloc(h, &h->part_nz, (size_t)IG_MAX_CANDS * h->grid_score * 25)) return -2;
        if (dev_alloc(h, &h->part_c, (size_t)IG_MAX_CANDS * h->grid_score * 2)) return -2;
        if (dev_alloc(h, &h->part_z, (size_t)IG_MAX_CANDS * h->grid_pre * 25)) return -2;
        if (dev_alloc(h, &h->part_i, (size_t)IG_MAX_CANDS * h->grid_pre * 25)) return -2;
        if (dev_alloc(h, &h->table, (size_t)IG_MAX_CANDS * IG_N_OPS * ns)) return -2;
        if (dev_alloc(h, &h->table_len, (size_t)IG_MAX_CANDS * IG_N_OPS * ns)) return -2;
        if (dev_alloc(h, &h->rowidx, (size_t)IG_MAX_CANDS * ns)) return -2;
        if (dev_alloc(h, &h->clstab, (size_t)IG_MAX_CANDS) || dev_alloc(h, &h->subx, (size_t)ns) || dev_alloc(h, &h->rinfo, (size_t)IG_MAX_CANDS * ns)) return -2;
        h->n_part_full = sms * 8;
        if (dev_alloc(h, &h->part_full, h->n_part_full)) return -2;
        if (dev_alloc(h, &h->part_out, h->n_part_full)) return -2;
        h->n_part_zc = sms * 2;
        if (dev_alloc(h, &h->part_zc, h->n_part_zc) || dev_alloc(h, &h->part_nc, h->n_part_zc)) return -2;
        if (dev_alloc(h, &h->d_perm, nf)) return -2;
        h->d_nuniq = h->sc->res_nuniq; h->d_nsub = h->sc->res_nsub;   // device addresses inside the result record
        if (dev_alloc(h, &h->d_hist, 1 << 16)) return -2;
        h->sym_diag = nullptr;
        CK(cudaMallocHost((void**)&h->h_sc, sizeof(DevScalars)));
        CK(cudaMallocHost((void**)&h->h_small, 64 * sizeof(int)));
        CK(cudaMemsetAsync(h->sc, 0, sizeof(DevScalars), h->stream));
        // uploads
        if (upload_state(h, data->frags13, h->live)) return -2;
        CK(cudaMemcpy(h->init_live, h->live, sizeof(FragRec) * nf, cudaMemcpyDeviceToDevice));
        {
            std::vector<SubRec> s(ns);
            for (int i = 0; i < ns; i++) {
                s[i].parent = data->sub_parent[i]; s[i].watson = data->sub_watson[i];
                s[i].crick = data->sub_crick[i]; s[i].j = data->sub_j[i];
                if (s[i].parent < 0 || s[i].parent >= nf) { h->err = "ig_create: sub_parent out of range"; return -1; }
            }
            CK(cudaMemcpy(h->sub, s.data(), sizeof(SubRec) * ns, cudaMemcpyHostToDevice));
        }
        if (data->row_ptr[0] != 0 || data->row_ptr[ns] != h->nnz) { h->err = "ig_create: row_ptr inconsistent with nnz"; return -1; }
        CK(cudaMemcpy(h->row_ptr, data->row_ptr, sizeof(long long) * ((size_t)ns + 1), cudaMemcpyHostToDevice));
        {
            const size_t chunk = 1 << 24;
            std::vector<int2> buf(std::min<size_t>(chunk, (size_t)h->nnz));
            for (size_t off = 0; off < (size_t)h->nnz; off += chunk) {
                const size_t n = std::min(chunk, (size_t)h->nnz - off);
                for (size_t i = 0; i < n; i++) { buf[i].x = data->col[off + i]; buf[i].y = data->val[off + i]; }
                CK(cudaMemcpy(h->cv + off, buf.data(), n * sizeof(int2), cudaMemcpyHostToDevice));
            }
        }
        CK(cudaMemcpy(h->init_prev, data->init_prev, sizeof(int) * nf, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->init_next, data->init_next, sizeof(int) * nf, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->orientable, data->orientable, sizeof(int) * nf, cudaMemcpyHostToDevice));
        // factorial table on the device (same libdevice calls as the reference's factorial())
        double* d16;
        CK(cudaMalloc((void**)&d16, 16 * sizeof(double)));
        k_init_tables<<<1, 32, 0, h->stream>>>(d16);
        double h16[16];
        CK(cudaMemcpyAsync(h16, d16, sizeof h16, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpyToSymbol(c_log10_fact, h16, sizeof h16));
        CK(cudaFree(d16));
        // label counter: labels in the initial scaffold are arbitrary; start above their maximum
        int maxlab = 0;
        for (int i = 0; i < nf; i++) maxlab = std::max(maxlab, data->frags13[(size_t)2 * nf + i]);
        CK(cudaMemcpy(&h->sc->max_label, &maxlab, sizeof(int), cudaMemcpyHostToDevice));
        return 0;
    };
    int rc = body();
    if (rc) { g_err = h->err; ig_destroy(h); return rc; }
    *out = h;
    return 0;
}

extern "C" void ig_destroy(ig_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    void* ptrs[] = {h->flat_cnt, h->flat_list, h->clstab, h->subx, h->rinfo, h->nb_ptr, h->nb_idx, h->nb_cdf, h->nb_nnz, h->cyc_frags, h->live, h->part_out, h->part_c, h->table, h->table_len, h->rowidx, h->init_live, h->sub, h->coord, h->clen, h->row_ptr, h->cv, h->sym_diag,
                    h->init_prev, h->init_next, h->orientable, h->sc, h->desc, h->exz, h->exz_test, h->chunk_cnt,
                    h->rows, h->row_cnt, h->part_nz, h->part_z, h->part_i, h->part_full, h->part_zc, h->part_nc,
                    h->d_perm, h->d_hist};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->h_sc) cudaFreeHost(h->h_sc);
    if (h->h_small) cudaFreeHost(h->h_small);
    for (int i = 0; i < 6; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 16; i++) if (h->evk[i]) cudaEventDestroy(h->evk[i]);
    for (int i = 0; i < 4; i++) for (int j = 0; j <= IG_MAX_CANDS; j++) if (h->graph[i][j]) cudaGraphExecDestroy(h->graph[i][j]);
    if (h->cyc_in) cudaFree(h->cyc_in);
    if (h->cyc_out) cudaFree(h->cyc_out);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->pf) cudaStreamDestroy(h->pf);
    if (h->ev_cuts) cudaEventDestroy(h->ev_cuts);
    if (h->ev_cls) cudaEventDestroy(h->ev_cls);
    if (h->ev_coords) cudaEventDestroy(h->ev_coords);
    if (h->ev_lnz) cudaEventDestroy(h->ev_lnz);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_sel) cudaEventDestroy(h->ev_sel);
    if (h->ev_out) cudaEventDestroy(h->ev_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

static int use(ig_handle* h) {
    if (!h) { g_err = "null handle"; return -1; }
    CK(cudaSetDevice(h->cfg.device));
    return 0;
}
static int launch_ok(ig_handle* h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { h->err = std::string(what) + ": " + cudaGetErrorString(e); return -2; }
    return 0;
}

extern "C" int ig_set_params(ig_handle* h, const float p8[8]) {
    if (use(h)) return -1;
    Params p; memcpy(&p, p8, sizeof p);
    k_set_params<<<1, 1, 0, h->stream>>>(h->sc, p, 0);
    k_exz_table<<<std::min(1024, (h->ns + 256) / 256), 256, 0, h->stream>>>(h->exz, h->ns + 1, h->sc, h->cfg.mean_sub_len_kb, 0);
    if (launch_ok(h, "set_params")) return -2;
    CK(cudaStreamSynchronize(h->stream));
    h->params_set = true;
    h->incr_valid = false;  // lnz_full / z_cur depend on the parameters
    return 0;
}

// renumber like modify_gl_cuda_buffer (CL:2715-2806): contigs listed in ascending index of their
// head fragment (sequential select_uniq_id_c), stable sort by length descending, id = NC-1-rank.
static void canonical_labels(int nf, int32_t* st13) {
    int32_t* pos = st13; int32_t* id_c = st13 + (size_t)2 * nf; int32_t* l_cont = st13 + (size_t)9 * nf;
    std::vector<int> heads;
    for (int i = 0; i < nf; i++) if (pos[i] == 0) heads.push_back(i);
    std::vector<int> order(heads.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return l_cont[heads[x]] > l_cont[heads[y]]; });
    std::vector<std::pair<int, int>> lut;  // (old label, rank)
    lut.reserve(heads.size());
    for (size_t r = 0; r < order.size(); r++) lut.push_back({id_c[heads[order[r]]], (int)r});
    std::stable_sort(lut.begin(), lut.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
    const int nc = (int)heads.size();
    for (int i = 0; i < nf; i++) {
        auto it = std::upper_bound(lut.begin(), lut.end(), std::make_pair(id_c[i], INT32_MAX));
        if (it != lut.begin() && (it - 1)->first == id_c[i]) id_c[i] = (nc - 1) - (it - 1)->second;
    }
}

extern "C" int ig_get_state(ig_handle* h, int32_t* out13) {
    if (use(h)) return -1;
    const int nf = h->nf;
    std::vector<FragRec> tmp(nf);
    CK(cudaMemcpyAsync(tmp.data(), h->live, sizeof(FragRec) * nf, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < nf; i++) {
        const int* v = reinterpret_cast<const int*>(&tmp[i].f);
        for (int k = 0; k < IG_N_FIELDS; k++) out13[(size_t)k * nf + i] = v[k];
    }
    canonical_labels(nf, out13);
    return 0;
}
extern "C" int ig_set_state(ig_handle* h, const int32_t* in13) {
    if (use(h)) return -1;
    if (upload_state(h, in13, h->live)) return -2;
    int maxlab = 0;
    for (int i = 0; i < h->nf; i++) maxlab = std::max(maxlab, in13[(size_t)2 * h->nf + i]);
    CK(cudaMemcpy(&h->sc->max_label, &maxlab, sizeof(int), cudaMemcpyHostToDevice));
    h->coords_fresh = false;
    h->incr_valid = false;
    return 0;
}
extern "C" int ig_get_valid_insert(ig_handle* h, int32_t out12[12]) {
    if (use(h)) return -1;
    CK(cudaMemcpy(out12, h->sc->valid, 12 * sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}
extern "C" int ig_set_valid_insert(ig_handle* h, const int32_t in12[12]) {
    if (use(h)) return -1;
    CK(cudaMemcpy(h->sc->valid, in12, 12 * sizeof(int), cudaMemcpyHostToDevice));
    return 0;
}
extern "C" int ig_bomb(ig_handle* h, const int32_t* perm) {
    if (use(h)) return -1;
    CK(cudaMemcpyAsync(h->d_perm, perm, sizeof(int) * h->nf, cudaMemcpyHostToDevice, h->stream));
    k_explode<<<(h->nf + 255) / 256, 256, 0, h->stream>>>(h->live, h->nf, h->d_perm);
    if (launch_ok(h, "explode")) return -2;
    int maxlab = h->nf;  // perm values are 0..NF-1
    CK(cudaMemcpyAsync(&h->sc->max_label, &maxlab, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->coords_fresh = false;
    h->incr_valid = false;
    return 0;
}

// head of step_sampler: fill_dist_single + eval_likelihood (CL:1407-1409)
static int refresh_current(ig_handle* h, cudaStream_t st, bool fork) {
    const float mbar = h->cfg.mean_sub_len_kb;
    if (fork) {  // side stream: starts after everything already queued on the main stream
        cudaEventRecord(h->ev_fork, h->stream);
        cudaStreamWaitEvent(st, h->ev_fork, 0);
    }
    k_coords<<<h->n_part_zc, IG_THREADS, 0, st>>>(h->live, h->sub, h->coord, h->clen, h->ns, h->sc, mbar, 0, h->part_zc, h->part_nc, 1, h->subx);
    k_reduce<<<1, 256, 0, st>>>(h->part_zc, h->n_part_zc, &h->sc->z_cur, h->part_nc, &h->sc->nintra_cur);
    if (fork) cudaEventRecord(h->ev_coords, st);
    if (h->profile) cudaEventRecord(h->ev[2], st);
    k_full_lnz<<<h->n_part_full, IG_THREADS, 0, st>>>(h->row_ptr, h->cv, h->coord, h->clen, h->ns, h->sc, mbar, 0, h->exz, h->part_full);
    if (h->profile) cudaEventRecord(h->ev[3], st);
    k_reduce<<<1, 256, 0, st>>>(h->part_full, h->n_part_full, &h->sc->lnz_full, nullptr, nullptr);
    if (fork) cudaEventRecord(h->ev_lnz, st);
    h->n_launches += 4;
    h->coords_fresh = true; h->coords_ever = true;
    return launch_ok(h, "refresh_current");
}

// k_score grid: the candidates of a step run side by side, so the resident capacity (3 CTAs per SM) is shared
// between them -- one wave for the whole step instead of one (mostly latency) wave per candidate
static int score_grid_x(const ig_handle* h, int n) {
    if (h->grid_split <= 0) return h->grid_score;
    return std::max(h->grid_score / std::max(n, 1), h->grid_score / 8) * h->grid_split;
}

static int score_candidates(ig_handle* h, int a, const int32_t* cands, int n, int first_flip_eject, bool overlap) {
    if (n <= 0 || n > IG_MAX_CANDS) { h->err = "n_cands out of range"; return -1; }
    if (a < 0 || a >= h->nf) { h->err = "id_frag out of range"; return -1; }
    for (int i = 0; i < n; i++) if (cands[i] < 0 || cands[i] >= h->nf) { h->err = "candidate out of range"; return -1; }
    const float mbar = h->cfg.mean_sub_len_kb;
    int* hs = h->h_small;
    hs[0] = n; hs[1] = a;
    for (int i = 0; i < IG_MAX_CANDS; i++) hs[2 + i] = i < n ? cands[i] : 0;
    CK(cudaMemcpyAsync(&h->sc->n_cands, hs, (2 + IG_MAX_CANDS) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    const FragRec* live = h->live;
    k_cand_setup<<<1, 32, 0, h->stream>>>(live, h->sc, h->desc, h->cfg.max_bounds_insert, first_flip_eject, nullptr);
    k_find_cuts<<<dim3((h->nf + 255) / 256, n), 256, 0, h->stream>>>(live, h->nf, h->sc, h->desc, h->clstab);
    k_classes<<<n, IG_N_OPS * 32, 0, h->stream>>>(h->sc, h->desc, h->clstab, h->rigid, h->cfg.mean_sub_len_kb);
    if (overlap) cudaStreamWaitEvent(h->stream, h->ev_coords, 0);
    if (h->rows_small) {
        k_rows_small<<<n, IG_ROW_CHUNK, 0, h->stream>>>(h->coord, h->ns, h->sc, h->n_chunks, h->rows, h->rowidx, h->row_cnt, h->ns,
                                                        h->clstab, h->row_ptr, h->rinfo);
    } else {
        k_rows_count<<<dim3(h->n_chunks, n), IG_ROW_CHUNK, 0, h->stream>>>(h->coord, h->ns, h->sc, h->chunk_cnt, h->n_chunks);
        k_rows_write<<<dim3(h->n_chunks, n), IG_ROW_CHUNK, 0, h->stream>>>(h->coord, h->ns, h->sc, h->chunk_cnt, h->n_chunks, h->rows,
                                                                           h->rowidx, h->row_cnt, h->ns, h->clstab, h->row_ptr, h->rinfo);
    }
    k_precompute<<<dim3(h->grid_pre, n), IG_PRE_THREADS, 0, h->stream>>>(h->coord, h->clen, live, h->sub, h->sc, h->desc, h->rows, h->ns,
                                                                    h->table, h->table_len, mbar, h->part_z, h->part_i);
    if (h->profile) cudaEventRecord(h->ev[4], h->stream);
    const int gsx = h->flat ? h->grid_flat : score_grid_x(h, n);
    if (h->flat) {
        k_pick<<<dim3(gsx, n), IG_THREADS, 0, h->stream>>>(h->cv, h->coord, h->clen, h->sc, h->desc, h->rowidx, h->ns, h->row_cnt,
                                                                    h->flat_cnt, h->chunk_stride, h->part_c, h->flat_list, h->flat_stride, mbar, h->exz,
                                                                    h->clstab, h->subx, h->rinfo);
        k_eval_flat<<<h->grid_score, IG_THREADS, IG_SCORE_SMEM, h->stream>>>(h->sc, h->desc, h->ns, h->flat_cnt, h->chunk_stride, h->flat_list,
                                                                            h->flat_stride, h->table, h->table_len, mbar, h->exz, h->part_nz, h->clstab, h->flat_items);
    } else
    k_score<<<dim3(gsx, n), IG_THREADS, IG_SCORE_SMEM, h->stream>>>(h->row_ptr, h->cv, h->coord, h->clen, h->sc, h->desc, h->rows, h->rowidx,
                                                                 h->ns, h->row_cnt, h->table, h->table_len, mbar, h->exz, h->part_nz,
                                                                 h->part_c, h->gs_div, h->clstab, h->subx, h->rinfo, h->sparse_div);
    if (h->profile) cudaEventRecord(h->ev[5], h->stream);
    h->n_launches += 7;
    if (overlap) cudaStreamWaitEvent(h->stream, h->ev_lnz, 0);
    k_finalize<<<n, 1024, 0, h->stream>>>(h->row_ptr, h->cv, h->coord, h->sc, h->desc, h->rows, h->rowidx, h->ns, h->row_cnt, h->table,
                                         h->table_len, mbar, h->exz, h->part_nz, h->part_c, h->flat ? h->grid_score : gsx, h->part_z, h->part_i,
                                         h->grid_pre, h->cfg.n_pix, h->cfg.compat_last_block, h->d_nuniq, h->d_nsub, 0, gsx, h->flat ? 1 : 0);
    return launch_ok(h, "score_candidates");
}

static int apply_and_post(ig_handle* h, int forced_cand, int forced_op) {
    const int nf = h->nf;
    k_apply<<<(nf + 255) / 256, 256, 0, h->stream>>>(h->live, nf, h->sc, h->desc, forced_cand, forced_op);
    k_post_scalars<<<1, 1, 0, h->stream>>>(h->sc, h->desc, forced_cand, forced_op);
    k_post<<<(nf + 255) / 256, 256, 0, h->stream>>>(h->live, nf, h->init_prev, h->init_next, h->orientable, h->sc, nullptr, nullptr, nullptr);
    h->coords_fresh = false;
    h->n_launches += 3;
    return launch_ok(h, "apply");
}

static int fetch_result(ig_handle* h, int n, const int32_t* cands, ig_step_result* out, bool applied, bool copy = true) {
    if (copy) {
        CK(cudaMemcpyAsync(h->h_sc, h->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    const DevScalars& s = *h->h_sc;
    memset(out, 0, sizeof *out);
    for (int i = 0; i < n * IG_N_OPS; i++) out->scores[i] = s.scores[i];
    out->lnz_full = s.lnz_full;
    for (int i = 0; i < n; i++) { out->n_uniq[i] = s.res_nuniq[i]; out->n_sub[i] = s.res_nsub[i]; }
    out->q4_hits = s.q4_hits;
    if (applied) {
        out->likelihood = s.likelihood;
        out->cand_index = s.win_cand; out->op_sampled = s.win_op;
        out->id_f_sampled = cands ? cands[s.win_cand] : -1;
        out->n_contigs = s.n_heads; out->sum_l_cont = s.sum_l_cont;
        out->dist = (3.0 * h->nf - 0.5 * (double)s.dist_half) / (3.0 * h->nf);
    }
    return 0;
}

// Everything one step enqueues (no host synchronisation): used directly and under graph capture.
//   full = 1: fill_dist_single + eval_likelihood over every contact (CL:1407-1409), on the side stream,
//             overlapped with the candidate setup on the main stream;
//   full = 0: incremental refresh from the previous step's mutation table (same values up to f64
//             summation order), O(rows of the last move).
static int enqueue_step(ig_handle* h, int full, int n_grid_cands, int cycle = 0) {
    const float mbar = h->cfg.mean_sub_len_kb;
    const int n = n_grid_cands;
    if (!cycle) cudaMemcpyAsync(&h->sc->n_cands, h->h_small, (2 + IG_MAX_CANDS) * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    cudaEventRecord(h->ev_fork, h->stream);
    cudaStreamWaitEvent(h->side, h->ev_fork, 0);
    cudaStreamWaitEvent(h->pf, h->ev_fork, 0);
    if (h->prefetch) {
        PfList L;
        int c = 0;
        auto add = [&](const void* ptr, size_t bytes) { L.p[c] = (const char*)ptr; L.n[c] = bytes; c++; };
        add(h->cv, sizeof(int2) * (size_t)h->nnz); add(h->row_ptr, sizeof(long long) * ((size_t)h->ns + 1));
        add(h->coord, sizeof(CoordRec) * (size_t)h->ns); add(h->clen, sizeof(int) * (size_t)h->ns);
        add(h->subx, sizeof(SubX) * (size_t)h->ns); add(h->sub, sizeof(SubRec) * (size_t)h->ns);
        add(h->live, sizeof(FragRec) * (size_t)h->nf); add(h->exz, sizeof(float) * ((size_t)h->ns + 1));
        add(h->init_prev, sizeof(int) * (size_t)h->nf); add(h->init_next, sizeof(int) * (size_t)h->nf);
        add(h->orientable, sizeof(int) * (size_t)h->nf);
        L.cnt = c;
        k_prefetch_l2<<<h->n_part_zc, 256, 0, h->pf>>>(L);
    }
    if (full) {
        k_coords<<<h->n_part_zc, IG_THREADS, 0, h->side>>>(h->live, h->sub, h->coord, h->clen, h->ns, h->sc, mbar, 0, h->part_zc,
                                                          h->part_nc, 1, h->subx);
        k_reduce<<<1, 256, 0, h->side>>>(h->part_zc, h->n_part_zc, &h->sc->z_cur, h->part_nc, &h->sc->nintra_cur);
        cudaEventRecord(h->ev_coords, h->side);
        if (h->profile && !h->capturing) cudaEventRecord(h->ev[2], h->side);
        k_full_lnz<<<h->n_part_full, IG_THREADS, 0, h->side>>>(h->row_ptr, h->cv, h->coord, h->clen, h->ns, h->sc, mbar, 0, h->exz,
                                                              h->part_full);
        if (h->profile && !h->capturing) cudaEventRecord(h->ev[3], h->side);
        k_reduce<<<1, 256, 0, h->side>>>(h->part_full, h->n_part_full, &h->sc->lnz_full, nullptr, nullptr);
        cudaEventRecord(h->ev_lnz, h->side);
    } else {
        k_commit_coords<<<std::min(h->n_part_zc, (h->ns + 255) / 256), 256, 0, h->side>>>(h->coord, h->clen, h->sc, h->rows, h->ns,
                                                                                         h->table, h->table_len, h->live, h->sub, h->subx);
        cudaEventRecord(h->ev_coords, h->side);
        cudaEventRecord(h->ev_lnz, h->side);
    }
    const FragRec* live = h->live;
#define IG_MARK(i) do { if (h->profile && !h->capturing) cudaEventRecord(h->evk[i], h->stream); } while (0)
    IG_MARK(0);
    k_cand_setup<<<1, 32, 0, h->stream>>>(live, h->sc, h->desc, h->cfg.max_bounds_insert, 1, cycle ? h->cyc_in : nullptr);
    IG_MARK(1);
    k_find_cuts<<<dim3((h->nf + 255) / 256, n), 256, 0, h->stream>>>(live, h->nf, h->sc, h->desc, h->clstab);
    cudaEventRecord(h->ev_cuts, h->stream);
    cudaStreamWaitEvent(h->pf, h->ev_cuts, 0);
    k_classes<<<n, IG_N_OPS * 32, 0, h->pf>>>(h->sc, h->desc, h->clstab, h->rigid, h->cfg.mean_sub_len_kb);   // beside the row list; k_score needs it
    cudaEventRecord(h->ev_cls, h->pf);
    cudaStreamWaitEvent(h->stream, h->ev_coords, 0);
    IG_MARK(2);
    if (h->rows_small) {
        IG_MARK(3);
        k_rows_small<<<n, IG_ROW_CHUNK, 0, h->stream>>>(h->coord, h->ns, h->sc, h->n_chunks, h->rows, h->rowidx, h->row_cnt, h->ns,
                                                        h->clstab, h->row_ptr, h->rinfo);
    } else {
        k_rows_count<<<dim3(h->n_chunks, n), IG_ROW_CHUNK, 0, h->stream>>>(h->coord, h->ns, h->sc, h->chunk_cnt, h->n_chunks);
        IG_MARK(3);
        k_rows_write<<<dim3(h->n_chunks, n), IG_ROW_CHUNK, 0, h->stream>>>(h->coord, h->ns, h->sc, h->chunk_cnt, h->n_chunks, h->rows,
                                                                           h->rowidx, h->row_cnt, h->ns, h->clstab, h->row_ptr, h->rinfo);
    }
    IG_MARK(4);
    k_precompute<<<dim3(h->grid_pre, n), IG_PRE_THREADS, 0, h->stream>>>(h->coord, h->clen, live, h->sub, h->sc, h->desc, h->rows, h->ns,
                                                                    h->table, h->table_len, mbar, h->part_z, h->part_i);
    if (h->profile && !h->capturing) cudaEventRecord(h->ev[4], h->stream);
    cudaStreamWaitEvent(h->stream, h->ev_cls, 0);
    IG_MARK(5);
    const int gsx = h->flat ? h->grid_flat : score_grid_x(h, n);
    if (h->flat) {
        k_pick<<<dim3(gsx, n), IG_THREADS, 0, h->stream>>>(h->cv, h->coord, h->clen, h->sc, h->desc, h->rowidx, h->ns, h->row_cnt,
                                                                    h->flat_cnt, h->chunk_stride, h->part_c, h->flat_list, h->flat_stride, mbar, h->exz,
                                                                    h->clstab, h->subx, h->rinfo);
        k_eval_flat<<<h->grid_score, IG_THREADS, IG_SCORE_SMEM, h->stream>>>(h->sc, h->desc, h->ns, h->flat_cnt, h->chunk_stride, h->flat_list,
                                                                            h->flat_stride, h->table, h->table_len, mbar, h->exz, h->part_nz, h->clstab, h->flat_items);
    } else
    k_score<<<dim3(gsx, n), IG_THREADS, IG_SCORE_SMEM, h->stream>>>(h->row_ptr, h->cv, h->coord, h->clen, h->sc, h->desc, h->rows, h->rowidx,
                                                                 h->ns, h->row_cnt, h->table, h->table_len, mbar, h->exz, h->part_nz,
                                                                 h->part_c, h->gs_div, h->clstab, h->subx, h->rinfo, h->sparse_div);
    if (h->profile && !h->capturing) cudaEventRecord(h->ev[5], h->stream);
    cudaStreamWaitEvent(h->stream, h->ev_lnz, 0);
    IG_MARK(6);
    k_finalize<<<n, 1024, 0, h->stream>>>(h->row_ptr, h->cv, h->coord, h->sc, h->desc, h->rows, h->rowidx, h->ns, h->row_cnt, h->table,
                                         h->table_len, mbar, h->exz, h->part_nz, h->part_c, h->flat ? h->grid_score : gsx, h->part_z, h->part_i,
                                         h->grid_pre, h->cfg.n_pix, h->cfg.compat_last_block, h->d_nuniq, h->d_nsub, 1, gsx, h->flat ? 1 : 0);
    IG_MARK(7);
    IG_MARK(8);
    // independent of apply/post: runs beside them on the side stream, joined before the result copy
    cudaEventRecord(h->ev_sel, h->stream);
    cudaStreamWaitEvent(h->side, h->ev_sel, 0);
    // (not with rigid pruning: there the contacts outside the slice windows keep their terms, as they do mathematically)
    if (!h->rigid)
        k_lnz_outside<<<h->grid_score, IG_THREADS, 0, h->side>>>(h->row_ptr, h->cv, h->coord, h->clen, h->sc, h->rows, h->rowidx, h->ns,
                                                                 h->table, h->table_len, mbar, h->exz, h->part_out);
    cudaEventRecord(h->ev_out, h->side);
    IG_MARK(9);
    k_apply<<<(h->nf + 255) / 256, 256, 0, h->stream>>>(h->live, h->nf, h->sc, h->desc, -1, -1);
    IG_MARK(10);
    k_post<<<(h->nf + 255) / 256, 256, 0, h->stream>>>(h->live, h->nf, h->init_prev, h->init_next, h->orientable, h->sc,
                                                       cycle ? h->cyc_out : nullptr, h->d_nuniq, h->d_nsub);
    IG_MARK(11);
    cudaStreamWaitEvent(h->stream, h->ev_out, 0);
    if (!cycle) {
        cudaMemcpyAsync(h->h_sc, h->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, h->stream);
    }
    return launch_ok(h, "enqueue_step");
}
#define IG_LAUNCHES_FULL 15
#define IG_LAUNCHES_INCR 12

static int get_graph(ig_handle* h, int full, cudaGraphExec_t* out, int cycle, int n) {
    cudaGraphExec_t& ge = h->graph[full + 2 * cycle][n];
    if (!ge && !h->graph_failed) {
        cudaGraph_t g = nullptr;
        h->capturing = true;
        cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            enqueue_step(h, full, n, cycle);
            e = cudaStreamEndCapture(h->stream, &g);
        }
        h->capturing = false;
        if (e == cudaSuccess && g) e = cudaGraphInstantiate(&ge, g, 0);
        if (g) cudaGraphDestroy(g);
        if (e != cudaSuccess || !ge) { ge = nullptr; h->graph_failed = true; cudaGetLastError(); }
    }
    *out = ge;
    return 0;
}

extern "C" int ig_step(ig_handle* h, int32_t id_frag, const int32_t* cands, int32_t n_cands, ig_step_result* out) {
    if (use(h)) return -1;
    if (!h->params_set) { h->err = "ig_step: parameters not set (ig_set_params)"; return -1; }
    if (!out) { h->err = "ig_step: null result"; return -1; }
    if (n_cands <= 0 || n_cands > IG_MAX_CANDS) { h->err = "n_cands out of range"; return -1; }
    if (id_frag < 0 || id_frag >= h->nf) { h->err = "id_frag out of range"; return -1; }
    for (int i = 0; i < n_cands; i++) if (cands[i] < 0 || cands[i] >= h->nf) { h->err = "candidate out of range"; return -1; }
    for (int i = 0; i < n_cands; i++) if (cands[i] == id_frag) { h->err = "candidate equals the visited fragment (reference quirk Q4; see DESIGN.md D1)"; return -1; }
    int* hs = h->h_small;
    hs[0] = n_cands; hs[1] = id_frag;
    for (int i = 0; i < IG_MAX_CANDS; i++) hs[2 + i] = i < n_cands ? cands[i] : 0;
    const int full = (!h->incr_valid || (h->refresh_every > 0 && h->steps_since_full >= h->refresh_every)) ? 1 : 0;
    cudaGraphExec_t ge = nullptr;
    if (h->use_graph && !h->profile) get_graph(h, full, &ge, 0, n_cands);
    cudaEventRecord(h->ev[0], h->stream);
    if (ge) {
        CK(cudaGraphLaunch(ge, h->stream));
    } else {
        if (enqueue_step(h, full, n_cands)) return -2;
    }
    cudaEventRecord(h->ev[1], h->stream);
    CK(cudaStreamSynchronize(h->stream));
    h->n_launches += (full ? IG_LAUNCHES_FULL : IG_LAUNCHES_INCR) - (h->rigid ? 1 : 0) + (h->prefetch ? 1 : 0) - (h->rows_small ? 1 : 0) + (h->flat ? 1 : 0);
    h->steps_since_full = full ? 1 : h->steps_since_full + 1;
    h->n_full += full;
    h->incr_valid = true;
    h->coords_fresh = false; h->coords_ever = true;
    int rc = fetch_result(h, n_cands, cands, out, true, false);
    if (rc) return rc;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->ms_step += ms;
    if (h->profile) {
        if (full && cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) == cudaSuccess) h->ms_full += ms;
        if (cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]) == cudaSuccess) h->ms_score += ms;
        for (int i = 0; i < 11; i++)
            if (cudaEventElapsedTime(&ms, h->evk[i], h->evk[i + 1]) == cudaSuccess) h->ms_k[i] += ms;
    }
    h->n_steps++;
    return 0;
}

// A whole MCMC cycle (or any run of steps) enqueued without a single host synchronisation in between:
// the host uploads the plan {fragment, sorted candidates} of every step (drawn with the reference's own
// RNG calls, which do not depend on the chain state), every step is one CUDA-graph replay that reads its
// plan entry and writes a compact record on the device, and the host synchronises once at the end.
// Semantically identical to n_steps calls of ig_step (IG.full_em inner loop, instagraal.py:217-241).
static int cycle_buffers(ig_handle* h, int n_steps) {
    if (n_steps > h->cyc_cap) {
        if (h->cyc_in) cudaFree(h->cyc_in);
        if (h->cyc_out) cudaFree(h->cyc_out);
        if (h->cyc_frags) cudaFree(h->cyc_frags);
        h->cyc_in = nullptr; h->cyc_out = nullptr; h->cyc_frags = nullptr; h->cyc_cap = 0;
        for (int i = 2; i < 4; i++) for (int j = 0; j <= IG_MAX_CANDS; j++) if (h->graph[i][j]) { cudaGraphExecDestroy(h->graph[i][j]); h->graph[i][j] = nullptr; }  // pointers are baked in
        if (dev_alloc(h, &h->cyc_in, (size_t)n_steps * (2 + IG_MAX_CANDS)) || dev_alloc(h, &h->cyc_out, (size_t)n_steps) ||
            dev_alloc(h, &h->cyc_frags, (size_t)n_steps)) return -2;
        h->cyc_cap = n_steps;
    }
    return 0;
}

// replay n_steps steps from the plan in h->cyc_in (already on the device, or being written by an earlier kernel of
// the stream); grid_n(t) = number of candidate slots the step's grid is built for
template <class GridN>
static int run_plan(ig_handle* h, int n_steps, GridN grid_n, ig_cycle_step* out) {
    CK(cudaMemsetAsync(&h->sc->step_idx, 0, sizeof(int), h->stream));
    cudaEventRecord(h->ev[0], h->stream);
    for (int t = 0; t < n_steps; t++) {
        const int full = (!h->incr_valid || (h->refresh_every > 0 && h->steps_since_full >= h->refresh_every)) ? 1 : 0;
        cudaGraphExec_t ge = nullptr;
        if (h->use_graph) get_graph(h, full, &ge, 1, grid_n(t));
        if (ge) { CK(cudaGraphLaunch(ge, h->stream)); }
        else if (enqueue_step(h, full, grid_n(t), 1)) return -2;
        h->steps_since_full = full ? 1 : h->steps_since_full + 1;
        h->n_full += full;
        h->incr_valid = true;
        h->n_launches += (full ? IG_LAUNCHES_FULL : IG_LAUNCHES_INCR) - (h->rigid ? 1 : 0) + (h->prefetch ? 1 : 0) - (h->rows_small ? 1 : 0) + (h->flat ? 1 : 0);
    }
    cudaEventRecord(h->ev[1], h->stream);
    std::vector<CycleOut> res(n_steps);
    std::vector<int> plan((size_t)n_steps * (2 + IG_MAX_CANDS));
    CK(cudaMemcpyAsync(res.data(), h->cyc_out, sizeof(CycleOut) * n_steps, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(plan.data(), h->cyc_in, plan.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->coords_fresh = false; h->coords_ever = true;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]) == cudaSuccess) h->ms_step += ms;
    h->n_steps += n_steps;
    for (int t = 0; t < n_steps; t++) {
        const CycleOut& r = res[t];
        const int* pl = &plan[(size_t)t * (2 + IG_MAX_CANDS)];
        ig_cycle_step& o = out[t];
        o.likelihood = r.likelihood; o.lnz_full = r.lnz_full;
        o.dist = (3.0 * h->nf - 0.5 * (double)r.dist_half) / (3.0 * h->nf);
        o.sum_l_cont = r.sum_l_cont; o.n_contigs = r.n_heads; o.op_sampled = r.win_op; o.cand_index = r.win_cand;
        o.id_f_sampled = pl[2 + r.win_cand];
        o.q4_hits = r.q4_hits; o.n_proposals = 0;
        for (int i = 0; i < pl[0]; i++) o.n_proposals += r.n_uniq[i];
    }
    return 0;
}

extern "C" int ig_run_cycle(ig_handle* h, int32_t n_steps, const int32_t* frags, const int32_t* cands8, const int32_t* n_cands,
                            ig_cycle_step* out) {
    if (use(h)) return -1;
    if (!h->params_set) { h->err = "ig_run_cycle: parameters not set"; return -1; }
    if (n_steps <= 0) return 0;
    std::vector<int> plan((size_t)n_steps * (2 + IG_MAX_CANDS), 0);
    for (int t = 0; t < n_steps; t++) {
        const int n = n_cands[t];
        if (n <= 0 || n > IG_MAX_CANDS || frags[t] < 0 || frags[t] >= h->nf) { h->err = "ig_run_cycle: bad plan entry"; return -1; }
        int* p = &plan[(size_t)t * (2 + IG_MAX_CANDS)];
        p[0] = n; p[1] = frags[t];
        for (int i = 0; i < n; i++) {
            const int c = cands8[(size_t)t * IG_MAX_CANDS + i];
            if (c < 0 || c >= h->nf || c == frags[t]) { h->err = "ig_run_cycle: candidate out of range or equal to the visited fragment"; return -1; }
            p[2 + i] = c;
        }
    }
    if (cycle_buffers(h, n_steps)) return -2;
    CK(cudaMemcpyAsync(h->cyc_in, plan.data(), plan.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    return run_plan(h, n_steps, [&](int t) { return n_cands[t]; }, out);
}

// setup_distri_frags (CL:3053-3101) on the device: per fragment the candidate fragments (self excluded), the
// running sum of their probabilities pk, and the number of non-zero pk (CL:3113).
extern "C" int ig_set_neighbour_weights(ig_handle* h, const int64_t* ptr, const int32_t* idx, const double* cdf, const int32_t* n_nonzero) {
    if (use(h)) return -1;
    if (!ptr || !n_nonzero) { h->err = "ig_set_neighbour_weights: null argument"; return -1; }
    const size_t n = (size_t)ptr[h->nf];
    for (void* q : {(void*)h->nb_ptr, (void*)h->nb_idx, (void*)h->nb_cdf, (void*)h->nb_nnz}) if (q) cudaFree(q);
    h->nb_ptr = nullptr; h->nb_idx = nullptr; h->nb_cdf = nullptr; h->nb_nnz = nullptr;
    if (dev_alloc(h, &h->nb_ptr, (size_t)h->nf + 1) || dev_alloc(h, &h->nb_idx, n + 1) || dev_alloc(h, &h->nb_cdf, n + 1) ||
        dev_alloc(h, &h->nb_nnz, (size_t)h->nf)) return -2;
    CK(cudaMemcpy(h->nb_ptr, ptr, sizeof(long long) * ((size_t)h->nf + 1), cudaMemcpyHostToDevice));
    if (n) {
        CK(cudaMemcpy(h->nb_idx, idx, sizeof(int) * n, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->nb_cdf, cdf, sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(h->nb_nnz, n_nonzero, sizeof(int) * (size_t)h->nf, cudaMemcpyHostToDevice));
    return 0;
}

// One sweep of step_sampler over `frags` with the neighbour draws made ON THE DEVICE (Philox4x32-10 keyed by
// (seed, cycle, step, draw)): the host uploads the visiting order, one kernel draws every step's candidates,
// then the steps replay without host synchronisation.  Needs ig_set_neighbour_weights.
extern "C" int ig_run_cycle_device(ig_handle* h, int32_t n_steps, const int32_t* frags, int32_t n_neighbours, uint64_t seed,
                                   uint32_t cycle, ig_cycle_step* out) {
    if (use(h)) return -1;
    if (!h->params_set) { h->err = "ig_run_cycle_device: parameters not set"; return -1; }
    if (!h->nb_ptr) { h->err = "ig_run_cycle_device: neighbour weights not set (ig_set_neighbour_weights)"; return -1; }
    if (n_neighbours <= 0 || n_neighbours > IG_MAX_CANDS) { h->err = "ig_run_cycle_device: n_neighbours out of range"; return -1; }
    if (h->nf < 2) { h->err = "ig_run_cycle_device: needs at least two fragments"; return -1; }
    if (n_steps <= 0) return 0;
    for (int t = 0; t < n_steps; t++) if (frags[t] < 0 || frags[t] >= h->nf) { h->err = "ig_run_cycle_device: fragment out of range"; return -1; }
    if (cycle_buffers(h, n_steps)) return -2;
    CK(cudaMemcpyAsync(h->cyc_frags, frags, sizeof(int) * (size_t)n_steps, cudaMemcpyHostToDevice, h->stream));
    k_draw_plan<<<(n_steps + 127) / 128, 128, 0, h->stream>>>(h->cyc_in, h->cyc_frags, n_steps, n_neighbours, h->nf, h->nb_ptr, h->nb_idx,
                                                            h->nb_cdf, h->nb_nnz, (unsigned)(seed & 0xffffffffu), (unsigned)(seed >> 32), cycle);
    if (launch_ok(h, "draw_plan")) return -2;
    h->n_launches += 1;
    return run_plan(h, n_steps, [&](int) { return (int)n_neighbours; }, out);
}

// download the plan of the last cycle ([n_steps][2 + IG_MAX_CANDS]: n_cands, fragment, candidates) -- tests / logging
extern "C" int ig_get_cycle_plan(ig_handle* h, int32_t n_steps, int32_t* out) {
    if (use(h)) return -1;
    if (n_steps > h->cyc_cap) { h->err = "ig_get_cycle_plan: no such plan"; return -1; }
    CK(cudaMemcpy(out, h->cyc_in, sizeof(int) * (size_t)n_steps * (2 + IG_MAX_CANDS), cudaMemcpyDeviceToHost));
    return 0;
}

// measurement / parity knobs: refresh_every = N -> recompute the coordinates and the full likelihood
// over every contact at least every N steps (1 = every step, exactly the reference's schedule;
// 0 = only when the state was changed from outside); use_graph = replay the step as one CUDA graph.
extern "C" int ig_set_options(ig_handle* h, int32_t refresh_every, int32_t use_graph) {
    if (!h) return -1;
    h->refresh_every = refresh_every;
    h->use_graph = use_graph ? true : false;
    return 0;
}

extern "C" int ig_eval_scores(ig_handle* h, int32_t id_frag, int32_t id_cand, int32_t flip_eject, double out24[24],
                              int32_t* n_uniq, int32_t* n_sub) {
    if (use(h)) return -1;
    if (!h->params_set) { h->err = "ig_eval_scores: parameters not set"; return -1; }
    h->incr_valid = false;
    if (refresh_current(h, h->stream, false)) return -2;
    if (int rc = score_candidates(h, id_frag, &id_cand, 1, flip_eject, false)) return rc;
    ig_step_result r;
    if (fetch_result(h, 1, &id_cand, &r, false)) return -2;
    for (int i = 0; i < 24; i++) out24[i] = r.scores[i];
    if (n_uniq) *n_uniq = r.n_uniq[0];
    if (n_sub) *n_sub = r.n_sub[0];
    return 0;
}

extern "C" int ig_apply(ig_handle* h, int32_t id_frag, int32_t id_cand, int32_t op, ig_step_result* out) {
    if (use(h)) return -1;
    if (op < 0 || op >= IG_N_OPS) { h->err = "ig_apply: bad op"; return -1; }
    h->incr_valid = false;
    if (id_frag < 0 || id_frag >= h->nf || id_cand < 0 || id_cand >= h->nf) { h->err = "ig_apply: fragment out of range"; return -1; }
    int* hs = h->h_small;
    hs[0] = 1; hs[1] = id_frag; hs[2] = id_cand;
    CK(cudaMemcpyAsync(&h->sc->n_cands, hs, 3 * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    int32_t saved[12];
    CK(cudaMemcpyAsync(saved, h->sc->valid, sizeof saved, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const FragRec* live = h->live;
    k_cand_setup<<<1, 32, 0, h->stream>>>(live, h->sc, h->desc, h->cfg.max_bounds_insert, 0, nullptr);
    k_find_cuts<<<dim3((h->nf + 255) / 256, 1), 256, 0, h->stream>>>(live, h->nf, h->sc, h->desc, h->clstab);
    // test_copy_struct only re-runs get_bounds for op >= 12 (CL:2121-2126): restore the list otherwise
    CK(cudaMemcpyAsync(h->sc->valid, saved, sizeof saved, cudaMemcpyHostToDevice, h->stream));
    if (apply_and_post(h, 0, op)) return -2;
    ig_step_result r;
    if (fetch_result(h, 1, &id_cand, &r, true)) return -2;
    r.op_sampled = op; r.id_f_sampled = id_cand; r.cand_index = 0;
    if (out) *out = r;
    return 0;
}

extern "C" int ig_full_likelihood(ig_handle* h, const float p8[8], int32_t use_stale_coords, double out3[3]) {
    if (use(h)) return -1;
    const float mbar = h->cfg.mean_sub_len_kb;
    Params p; memcpy(&p, p8, sizeof p);
    k_set_params<<<1, 1, 0, h->stream>>>(h->sc, p, 1);
    k_exz_table<<<std::min(1024, (h->ns + 256) / 256), 256, 0, h->stream>>>(h->exz_test, h->ns + 1, h->sc, mbar, 1);
    const int write = (use_stale_coords && h->coords_ever) ? 0 : 1;
    if (write) { h->coords_ever = true; h->coords_fresh = true; }
    k_coords<<<h->n_part_zc, IG_THREADS, 0, h->stream>>>(h->live, h->sub, h->coord, h->clen, h->ns, h->sc, mbar, 1,
                                                        h->part_zc, h->part_nc, write, h->subx);
    k_reduce<<<1, 256, 0, h->stream>>>(h->part_zc, h->n_part_zc, &h->sc->full_out[1], h->part_nc, &h->sc->full_nintra);
    k_full_lnz<<<h->n_part_full, IG_THREADS, 0, h->stream>>>(h->row_ptr, h->cv, h->coord, h->clen, h->ns, h->sc, mbar, 1,
                                                            h->exz_test, h->part_full);
    k_reduce<<<1, 256, 0, h->stream>>>(h->part_full, h->n_part_full, &h->sc->full_out[0], nullptr, nullptr);
    if (launch_ok(h, "full_likelihood")) return -2;
    CK(cudaMemcpyAsync(h->h_sc, h->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    out3[0] = h->h_sc->full_out[0];
    out3[1] = h->h_sc->full_out[1];
    out3[2] = (double)h->h_sc->full_nintra;
    return 0;
}

extern "C" int ig_distance_histogram(ig_handle* h, double bin_kb, double max_kb, int32_t n_rows, int32_t n_bins,
                                     int64_t* hist, int64_t* rows_used) {
    if (use(h)) return -1;
    if (n_bins <= 0 || n_bins > (1 << 16) - 1) { h->err = "ig_distance_histogram: n_bins out of range"; return -1; }
    if (n_rows > h->ns) n_rows = h->ns;
    CK(cudaMemsetAsync(h->d_hist, 0, sizeof(unsigned long long) * (n_bins + 1), h->stream));
    k_histogram<<<h->n_part_full, IG_THREADS, 0, h->stream>>>(h->row_ptr, h->cv, h->sym_diag, h->init_live, h->sub, n_rows, bin_kb,
                                                             max_kb, n_bins, h->d_hist, h->d_hist + n_bins);
    if (launch_ok(h, "histogram")) return -2;
    std::vector<unsigned long long> tmp(n_bins + 1);
    CK(cudaMemcpyAsync(tmp.data(), h->d_hist, sizeof(unsigned long long) * (n_bins + 1), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < n_bins; i++) hist[i] = (int64_t)tmp[i];
    *rows_used = (int64_t)tmp[n_bins];
    return 0;
}

extern "C" int ig_set_sym_diag(ig_handle* h, const int32_t* diag) {
    if (use(h)) return -1;
    if (!h->sym_diag) { if (dev_alloc(h, &h->sym_diag, h->ns)) return -2; }
    CK(cudaMemcpy(h->sym_diag, diag, sizeof(int) * h->ns, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int ig_device_state_ptr(ig_handle* h, void** dev_ptr, int64_t* n_bytes) {
    if (use(h)) return -1;
    CK(cudaStreamSynchronize(h->stream));
    *dev_ptr = h->live;
    *n_bytes = (int64_t)sizeof(FragRec) * h->nf;
    return 0;
}

// measurement: device time (CUDA events on the handle's stream) and algorithmic traffic counters.
// out[0..2] = ms in ig_step total / in k_score / in k_full_lnz (the latter two only while
// profiling is on); out[3] = kernel launches; out[4] = steps; out[5..9] = contacts read, rows,
// fragments, contacts selected, proposals scored by the scoring kernel.
extern "C" int ig_get_stats(ig_handle* h, double out10[10], int32_t reset) {
    if (use(h)) return -1;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(h->h_sc, h->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost));
    out10[0] = h->ms_step; out10[1] = h->ms_score; out10[2] = h->ms_full;
    out10[3] = (double)h->n_launches; out10[4] = (double)h->n_steps;
    h->last_n_full = h->n_full;
    out10[5] = (double)h->h_sc->st_contacts; out10[6] = (double)h->h_sc->st_rows; out10[7] = (double)h->h_sc->st_frags;
    out10[8] = (double)h->h_sc->st_selected; out10[9] = (double)h->h_sc->st_proposals;
    if (reset) {
        h->ms_step = h->ms_score = h->ms_full = 0.0; h->n_launches = 0; h->n_steps = 0; h->n_full = 0;
        CK(cudaMemset(&h->sc->st_contacts, 0, 5 * sizeof(unsigned long long)));
    }
    return 0;
}
extern "C" int ig_set_profiling(ig_handle* h, int32_t on) {
    if (!h) return -1;
    h->profile = on ? 1 : 0;
    return 0;
}

extern "C" int ig_get_full_refresh_count(ig_handle* h, int64_t* out) {
    if (!h) return -1;
    *out = h->last_n_full;
    return 0;
}

// profiling mode only: accumulated ms between consecutive main-stream launches of a step, in order:
// cand_setup, find_cuts(+wait coords), rows_count, rows_write, precompute, score(+wait lnz), finalize,
// select, lnz_outside, apply, post
extern "C" int ig_get_kernel_times(ig_handle* h, double out11[11], int32_t reset) {
    if (!h) return -1;
    for (int i = 0; i < 11; i++) { out11[i] = h->ms_k[i]; if (reset) h->ms_k[i] = 0.0; }
    return 0;
}

// display_current_matrix (CL:2555-2606) without the NS x NS host densification: sub_rank[s] = position of
// sub-fragment s in the displayed order (computed by the caller from ig_get_state like CL:2563-2585);
// out = K*K uint32 symmetric binned contact counts of the strict upper triangle.
extern "C" int ig_contact_thumbnail(ig_handle* h, const int32_t* sub_rank, int32_t K, uint32_t* out) {
    if (use(h)) return -1;
    if (K <= 0 || K > 8192) { h->err = "ig_contact_thumbnail: K out of range"; return -1; }
    int* d_rank = nullptr; unsigned int* d_img = nullptr;
    if (dev_alloc(h, &d_rank, h->ns) || dev_alloc(h, &d_img, (size_t)K * K)) return -2;
    CK(cudaMemcpyAsync(d_rank, sub_rank, sizeof(int) * h->ns, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(d_img, 0, sizeof(unsigned int) * (size_t)K * K, h->stream));
    k_thumbnail<<<h->n_part_full, IG_THREADS, 0, h->stream>>>(h->row_ptr, h->cv, d_rank, h->ns, K, d_img);
    if (launch_ok(h, "thumbnail")) return -2;
    CK(cudaMemcpyAsync(out, d_img, sizeof(unsigned int) * (size_t)K * K, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d_rank); cudaFree(d_img);
    return 0;
}

// timeline of the last cycle run (IG_TIMELINE builds only): out[n_steps][IG_TL_KERNELS][2] = earliest block start /
// latest block end of each kernel in %globaltimer ns (~0 / 0 when the kernel did not run in that step)
extern "C" int ig_timeline_reset(ig_handle* h) {
    if (use(h)) return -1;
#ifdef IG_TIMELINE
    std::vector<unsigned long long> init((size_t)IG_TL_STEPS * IG_TL_KERNELS * 2);
    for (size_t i = 0; i < init.size(); i += 2) { init[i] = ~0ull; init[i + 1] = 0ull; }
    CK(cudaMemcpyToSymbol(g_tl, init.data(), init.size() * sizeof(unsigned long long)));
    int zero = 0;
    CK(cudaMemcpyToSymbol(g_tl_step, &zero, sizeof zero));
    return 0;
#else
    h->err = "library built without -DIG_TIMELINE";
    return -1;
#endif
}
extern "C" int ig_timeline_blocks(ig_handle* h, int32_t n_blocks, uint64_t* out) {
    if (use(h)) return -1;
#ifdef IG_TIMELINE
    if (n_blocks > IG_TL_BLOCKS) n_blocks = IG_TL_BLOCKS;
    CK(cudaMemcpyFromSymbol(out, g_tlb, (size_t)n_blocks * 4 * sizeof(unsigned long long)));
    return 0;
#else
    (void)n_blocks; (void)out;
    h->err = "library built without -DIG_TIMELINE";
    return -1;
#endif
}
extern "C" int ig_timeline_phases(ig_handle* h, uint64_t* out8, int32_t reset) {
    if (use(h)) return -1;
#ifdef IG_TIMELINE
    CK(cudaMemcpyFromSymbol(out8, g_tlp, 16 * sizeof(unsigned long long)));
    if (reset) { unsigned long long z[16] = {0}; CK(cudaMemcpyToSymbol(g_tlp, z, sizeof z)); }
    return 0;
#else
    (void)out8; (void)reset;
    h->err = "library built without -DIG_TIMELINE";
    return -1;
#endif
}
extern "C" int ig_timeline_get(ig_handle* h, int32_t n_steps, uint64_t* out) {
    if (use(h)) return -1;
#ifdef IG_TIMELINE
    if (n_steps > IG_TL_STEPS) n_steps = IG_TL_STEPS;
    CK(cudaMemcpyFromSymbol(out, g_tl, (size_t)n_steps * IG_TL_KERNELS * 2 * sizeof(unsigned long long)));
    return 0;
#else
    (void)n_steps; (void)out;
    h->err = "library built without -DIG_TIMELINE";
    return -1;
#endif
}

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
    double obc_total;   // sum over every stored contact of the observed-count-only part of its term (k_obc_sum)
    unsigned int ticket_cuts[IG_MAX_CANDS], ticket_rows[IG_MAX_CANDS], ticket_fin;  // last-block-done counters
    // measurement: algorithmic traffic of the scoring kernel, accumulated over steps
    unsigned long long st_contacts, st_rows, st_frags, st_selected, st_proposals;
    // flat scoring path (small levels): per candidate, list slots owned by the affected rows (padded row lengths)
    int flat_segtotal[IG_MAX_CANDS];
    // per-candidate counters of the step (written by k_finalize): they travel with this record in ONE copy
    int res_nuniq[IG_MAX_CANDS], res_nsub[IG_MAX_CANDS];
    // streaming scoring path of large levels: which candidates take it, overflow of a pick list
    int use_stream[IG_MAX_CANDS];
    int list_overflow, pad2_;
};

struct CycleOut {  // compact per-step record of ig_run_cycle (128 B)
    double likelihood, lnz_full;
    long long dist_half, sum_l_cont;
    int n_heads, win_cand, win_op, q4_hits;
    int n_uniq[IG_MAX_CANDS], n_sub[IG_MAX_CANDS];
    int pad[8];
};

#include "ig_k_common.cuh"   // optional device-side timeline, the reference's device math (textual twins), reduction helpers
#include "ig_k_state.cuh"   // state of the current scaffold: coordinates, full likelihood over every contact, tables
#include "ig_k_setup.cuh"   // per-candidate setup: descriptor, cut fragments, rigid-motion class tables, ordered affected-row list
#include "ig_k_score.cuh"   // scoring: slice rule, row-end table + zero terms, evaluation queue, the row-per-warp scoring kernel
#include "ig_k_flat.cuh"   // flat scoring path of small levels: k_pick + k_eval_flat
#include "ig_k_finish.cuh"   // finalisation, move selection, incremental maintenance, apply, contig bookkeeping, histogram, thumbnail
#include "ig_k_rng.cuh"   // production RNG mode: Philox4x32-10 neighbour draws of a whole cycle

// ================================================================================================
// host side
// the level's read-only arrays: shared by every chain cloned from one handle (ig_clone), freed with the last of them
struct LevelBlock {
    int refs;
    int max_val;   // largest observed count of the level (decides whether the packed likelihood records can hold it)
    double val_total;   // sum of the observed counts of every stored contact
    long long max_row_len;   // longest CSR row (sizes the pick list of the streaming scoring path)
    int* sym_diag;
    long long* nb_ptr; int* nb_idx; double* nb_cdf; int* nb_nnz;
};
struct ig_replica_state;
struct ig_handle {
    ig_config cfg;
    LevelBlock* lvl;
    ig_replica_state* rep;   // NCCL communicator + gather buffers (ig_nccl_init), owned by the process's lead handle
    int nf, ns;
    long long nnz;
    cudaStream_t stream, side, pf;
    cudaEvent_t ev_coords, ev_lnz, ev_fork, ev_sel, ev_out, ev_cuts, ev_cls, ev_pre, ev_ksc;
    bool rows_small;  // affected-row list in one launch (small levels)
    bool flat;        // flat scoring path (k_pick + k_eval_flat) for small levels
    int grid_flat, flat_items; int* flat_cnt; FlatRec* flat_list; size_t flat_stride, chunk_stride;
    bool prefetch;  // the level's arrays fit the L2 comfortably: prefetch them at the start of a step
    bool streaming; // streaming scoring path (k_stream + k_eval_flat<true>): large levels with rigid_pruning != 0
    unsigned* bitmap; int bitmap_words; unsigned short* cls16; FlatRec* pick_list; size_t pick_cap; int stream_smem;
    int stream_rows_max;   // candidates with at most this many affected rows take the streaming path (exact mode: bounded by the pick list)
    FragRec *live, *init_live;
    SubRec* sub;
    CoordRec* coord;
    int* clen;
    long long* row_ptr;
    int2* cv;
    int *init_prev, *init_next, *orientable;
    DevScalars* sc;
    IgDescriptor* desc;
    float *exz, *exz_test;
    int n_chunks, *chunk_cnt, *rows, *row_cnt;
    int grid_score, grid_score_full;   // scoring grid in use / the whole machine (ig_set_gpu_share)
    double *part_nz, *part_z; int *part_i, *part_c;
    int grid_pre;
    RowMut* table; int *table_len, *rowidx;
    IgClassTab* clstab; int rigid; SubX* subx; RowInfo* rinfo;
    int* cyc_frags;
    double *part_full; int n_part_full;
    // cached per-contact records of the full likelihood (k_lnz_refresh / k_lnz_stream): 8 B x nnz per chain
    cudaGraphExec_t graph_nuis = nullptr; bool nuis_graph_failed = false; float *h_p8 = nullptr, *d_p8 = nullptr;   // nuisance evaluation as one graph
    int2* lnz_rec; unsigned char* row_dirty; int dp_bits; bool lnz_cache; int grid_lnz; long long lnz_pairs; int lnz_pad;
    double *part_zc; int* part_nc; int n_part_zc;
    int *d_nuniq, *d_nsub, *d_perm;
    unsigned long long* d_hist;
    DevScalars* h_sc;  // pinned mirror
    int* h_small;      // pinned scratch (cands, nuniq, nsub)
    bool params_set, coords_fresh, coords_ever; int init_max_label;
    long long label_hi = 0;          // upper bound of the contig labels in use (2 new ones per applied move), see labels_guard
    bool incr_valid; int refresh_every; long long steps_since_full;
    bool vinter_pos;  // live v_inter > 0 (the per-contact records assume it)
    bool lnz_stale;   // the parameters changed since lnz_full / z_cur were computed (the scaffold bookkeeping is intact)
    double* part_out;
    int gs_div, sparse_div, grid_split;
    long long last_n_full;
    cudaGraphExec_t graph[6][IG_MAX_CANDS + 1];   // [step kind + 3 * cycle][candidates in the grid]
    int* cyc_in; CycleOut* cyc_out; int cyc_cap; bool graph_failed, capturing, use_graph; long long n_full;
    int pending_steps;   // steps enqueued by an asynchronous cycle call and not collected yet
    // measurement (CUDA events on the launching stream)
    cudaEvent_t ev[6];
    cudaEvent_t evk[16]; double ms_k[16];  // per-kernel event timing of the main stream (profiling mode)
    double ms_step, ms_score, ms_full, ms_nuis; long long n_nuis;
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
// Host-to-device copy from pageable memory that is COMPLETE when it returns: a plain cudaMemcpy may return once the data
// is staged, and the kernels run on non-blocking streams that are not ordered against the legacy stream it uses.
static inline cudaError_t h2d_sync(void* dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(cudaStreamLegacy);
}

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

// parent != nullptr: a further chain on the same level (ig_clone): the level's read-only arrays are borrowed, not uploaded
static int create_impl(const ig_config* cfg, const ig_level_data* data, ig_handle* parent, ig_handle** out) {
    ig_handle* h = nullptr;
    if (!cfg || (!data && !parent) || !out) { g_err = "ig_create: null argument"; return -1; }
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
    h->incr_valid = false; h->lnz_stale = false; h->refresh_every = 4096; h->steps_since_full = 0;
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
    h->graph_failed = false; h->capturing = false; h->use_graph = true; h->n_full = 0; h->pending_steps = 0; h->rep = nullptr;
    cudaError_t e0 = cudaSetDevice(cfg->device);
    if (e0 != cudaSuccess) { g_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e0); delete h; return -2; }
#define CKC(x) do { int r_ = (x); if (r_) { g_err = h->err; ig_destroy(h); return r_; } } while (0)
    auto body = [&]() -> int {
        CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&h->pf, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_cuts, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_cls, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_pre, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_ksc, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_coords, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_lnz, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_sel, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
        for (int i = 0; i < 6; i++) CK(cudaEventCreate(&h->ev[i]));
        for (int i = 0; i < 16; i++) { CK(cudaEventCreate(&h->evk[i])); h->ms_k[i] = 0.0; }
        h->ms_step = h->ms_score = h->ms_full = 0.0; h->ms_nuis = 0.0; h->n_nuis = 0; h->n_launches = 0; h->n_steps = 0; h->profile = 0;
        const int nf = h->nf, ns = h->ns;
        if (dev_alloc(h, &h->live, nf) || dev_alloc(h, &h->coord, ns) || dev_alloc(h, &h->clen, ns)) return -2;
        if (parent) {
            h->lvl = parent->lvl; h->lvl->refs++;
            h->init_live = parent->init_live; h->sub = parent->sub; h->row_ptr = parent->row_ptr; h->cv = parent->cv;
            h->init_prev = parent->init_prev; h->init_next = parent->init_next; h->orientable = parent->orientable;
        } else {
            h->lvl = new LevelBlock();
            memset(h->lvl, 0, sizeof(LevelBlock)); h->lvl->refs = 1;
            if (dev_alloc(h, &h->init_live, nf) || dev_alloc(h, &h->sub, ns)) return -2;
            if (dev_alloc(h, &h->row_ptr, (size_t)ns + 1) || dev_alloc(h, &h->cv, (size_t)h->nnz + 2)) return -2;
            if (dev_alloc(h, &h->init_prev, nf) || dev_alloc(h, &h->init_next, nf) || dev_alloc(h, &h->orientable, nf)) return -2;
        }
        if (dev_alloc(h, &h->sc, 1) || dev_alloc(h, &h->desc, IG_MAX_CANDS)) return -2;
        if (dev_alloc(h, &h->exz, (size_t)ns + 3) || dev_alloc(h, &h->exz_test, (size_t)ns + 3)) return -2;   // + the two slots of k_lnz_stream
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
        h->bitmap = nullptr; h->cls16 = nullptr; h->pick_list = nullptr;
        h->bitmap_words = (ns + 31) / 32;
        h->stream_smem = h->bitmap_words * (int)sizeof(unsigned);
        // rigid pruning: every linear candidate (a few per cent of its contacts are picked).  Reference-faithful mode: almost
        // every selected contact is a pick, so only candidates whose contacts are sure to fit the pick list take it (the
        // mid-assembly regime: two contigs of a few hundred sub-fragments, where the row-per-warp kernel is all latency)
        h->streaming = !h->flat && h->stream_smem <= 200 * 1024;
        if (const char* e = getenv("IG_STREAM")) h->streaming = h->streaming && atoi(e) != 0;   // experiments / tests
        if (getenv("IG_FORCE_STREAM") && h->stream_smem <= 200 * 1024) {   // tests: small levels through the streaming path
            h->streaming = true; h->flat = false;
        }
        h->stream_rows_max = 0;
        if (h->streaming) {
            h->rows_small = false;   // the two-pass row list also writes the bitmap and the class words
            h->pick_cap = std::min<size_t>((size_t)h->nnz + 64, (size_t)2 << 20);   // 64-byte records: 128 MB per candidate slot at most
            if (const char* e = getenv("IG_PICK_CAP")) h->pick_cap = (size_t)std::max(64, atoi(e));
            if (!parent) {
                long long mr = 0;
                for (int r = 0; r < ns; r++) mr = std::max(mr, (long long)(data->row_ptr[r + 1] - data->row_ptr[r]));
                h->lvl->max_row_len = mr;
            }
            h->stream_rows_max = h->rigid ? INT32_MAX : (int)std::min<long long>(INT32_MAX, (long long)h->pick_cap / std::max<long long>(h->lvl->max_row_len, 1));
            if (const char* e = getenv("IG_STREAM_ROWS")) h->stream_rows_max = atoi(e);   // experiments
            if (dev_alloc(h, &h->bitmap, (size_t)IG_MAX_CANDS * h->bitmap_words) || dev_alloc(h, &h->cls16, (size_t)IG_MAX_CANDS * ns) ||
                dev_alloc(h, &h->pick_list, (size_t)IG_MAX_CANDS * h->pick_cap)) return -2;
            CK(cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, h->stream_smem));
        }
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, cfg->device));
        const int sms = prop.multiProcessorCount;
        h->grid_score = sms * IG_SCORE_CTAS_PER_SM;  // 24 resident warps per SM (80 registers per thread; 6 KB of accumulators per warp)
        h->grid_score_full = h->grid_score;
        CK(cudaFuncSetAttribute(k_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IG_SCORE_SMEM));
        CK(cudaFuncSetAttribute(k_eval_flat<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IG_SCORE_SMEM));
        CK(cudaFuncSetAttribute(k_eval_flat<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)IG_SCORE_SMEM));
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
        if (dev_alloc(h, &h->part_full, 2 * (size_t)h->n_part_full)) return -2;
        h->lnz_rec = nullptr; h->row_dirty = nullptr; h->lnz_cache = false; h->grid_lnz = sms * 4;
        if (dev_alloc(h, &h->part_out, h->n_part_full)) return -2;
        h->n_part_zc = sms * 2;
        if (dev_alloc(h, &h->part_zc, h->n_part_zc) || dev_alloc(h, &h->part_nc, h->n_part_zc)) return -2;
        if (dev_alloc(h, &h->d_perm, nf)) return -2;
        h->d_nuniq = h->sc->res_nuniq; h->d_nsub = h->sc->res_nsub;   // device addresses inside the result record
        if (dev_alloc(h, &h->d_hist, 1 << 16)) return -2;
        CK(cudaMallocHost((void**)&h->h_sc, sizeof(DevScalars)));
        CK(cudaMallocHost((void**)&h->h_small, 64 * sizeof(int)));
        CK(cudaMallocHost((void**)&h->h_p8, sizeof(Params)));
        if (dev_alloc(h, &h->d_p8, 8)) return -2;
        CK(cudaMemsetAsync(h->sc, 0, sizeof(DevScalars), h->stream));
        // uploads
        if (parent) {
            CK(cudaMemcpyAsync(h->live, h->init_live, sizeof(FragRec) * nf, cudaMemcpyDeviceToDevice, h->stream));
        } else {
            if (upload_state(h, data->frags13, h->live)) return -2;
            CK(cudaMemcpy(h->init_live, h->live, sizeof(FragRec) * nf, cudaMemcpyDeviceToDevice));
            {
                std::vector<SubRec> s(ns);
                for (int i = 0; i < ns; i++) {
                    s[i].parent = data->sub_parent[i]; s[i].watson = data->sub_watson[i];
                    s[i].crick = data->sub_crick[i]; s[i].j = data->sub_j[i];
                    if (s[i].parent < 0 || s[i].parent >= nf) { h->err = "ig_create: sub_parent out of range"; return -1; }
                }
                CK(h2d_sync(h->sub, s.data(), sizeof(SubRec) * ns));
            }
            if (data->row_ptr[0] != 0 || data->row_ptr[ns] != h->nnz) { h->err = "ig_create: row_ptr inconsistent with nnz"; return -1; }
            CK(h2d_sync(h->row_ptr, data->row_ptr, sizeof(long long) * ((size_t)ns + 1)));
            {
                const size_t chunk = 1 << 24;
                std::vector<int2> buf(std::min<size_t>(chunk, (size_t)h->nnz));
                for (size_t off = 0; off < (size_t)h->nnz; off += chunk) {
                    const size_t n = std::min(chunk, (size_t)h->nnz - off);
                    for (size_t i = 0; i < n; i++) { buf[i].x = data->col[off + i]; buf[i].y = data->val[off + i]; }
                    CK(h2d_sync(h->cv + off, buf.data(), n * sizeof(int2)));
                }
            }
            CK(cudaMemset(h->cv + h->nnz, 0, 2 * sizeof(int2)));   // k_full_lnz reads contacts in aligned pairs
            CK(h2d_sync(h->init_prev, data->init_prev, sizeof(int) * nf));
            CK(h2d_sync(h->init_next, data->init_next, sizeof(int) * nf));
            CK(h2d_sync(h->orientable, data->orientable, sizeof(int) * nf));
        }
        if (!parent) {
            int mv = 0; long long tot = 0;
            for (long long i = 0; i < h->nnz; i++) { mv = std::max(mv, data->val[i]); tot += data->val[i]; }
            h->lvl->max_val = mv; h->lvl->val_total = (double)tot;
        }
        {   // packed likelihood records: table index (sub-fragment separation, or one of two special slots) in the low
            // bits, the observed count above them; streamed in whole trips of 64 record pairs
            h->dp_bits = 1;
            while ((1 << h->dp_bits) < ns + 3) h->dp_bits++;
            h->lnz_pairs = ((h->nnz + 1) / 2 + 63) / 64 * 64;
            h->lnz_pad = (int)(2 * h->lnz_pairs - h->nnz);
            h->lnz_cache = h->nnz > 0 && h->dp_bits <= 24 && (long long)h->lvl->max_val < (1LL << (32 - h->dp_bits));
            if (const char* e = getenv("IG_LNZ_CACHE")) h->lnz_cache = h->lnz_cache && atoi(e) != 0;   // tests / A-B runs
            if (h->lnz_cache) {
                if (dev_alloc(h, &h->lnz_rec, (size_t)(2 * h->lnz_pairs)) || dev_alloc(h, &h->row_dirty, (size_t)ns)) return -2;
                std::vector<int2> pad(h->lnz_pad);
                for (auto& q : pad) { const float m1 = -1.0f; memcpy(&q.x, &m1, 4); q.y = ns + 2; }   // "skip" records
                if (h->lnz_pad) CK(h2d_sync(h->lnz_rec + h->nnz, pad.data(), sizeof(int2) * pad.size()));
                CK(cudaMemsetAsync(h->row_dirty, 1, (size_t)ns, h->stream));   // every row's records are built on first use
            }
        }
        // factorial table on the device (same libdevice calls as the reference's factorial())
        double* d16;
        CK(cudaMalloc((void**)&d16, 16 * sizeof(double)));
        k_init_tables<<<1, 32, 0, h->stream>>>(d16);
        double h16[16];
        CK(cudaMemcpyAsync(h16, d16, sizeof h16, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpyToSymbol(c_log10_fact, h16, sizeof h16));
        CK(cudaFree(d16));
        {   // log10_f32: bucket centres of the float mantissa
            double2 tab[128];
            for (int i = 0; i < 128; i++) { const double c = 1.0 + (i + 0.5) / 128.0; tab[i].x = 1.0 / c; tab[i].y = log10(c); }
            CK(cudaMemcpyToSymbol(g_log10tab, tab, sizeof tab));
        }
        // observed-count-only part of the likelihood terms, summed once (it depends on neither scaffold nor parameters)
        if (parent) {
            CK(cudaMemcpyAsync(&h->sc->obc_total, &parent->sc->obc_total, sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        } else {
            k_obc_sum<<<h->n_part_full, IG_THREADS, 0, h->stream>>>(h->cv, h->nnz, h->part_full);
            k_reduce<<<1, 256, 0, h->stream>>>(h->part_full, h->n_part_full, &h->sc->obc_total, nullptr, nullptr);
        }
        CK(cudaStreamSynchronize(h->stream));
        // label counter: labels in the initial scaffold are arbitrary; start above their maximum
        int maxlab = 0;
        if (parent) maxlab = parent->init_max_label;
        else for (int i = 0; i < nf; i++) maxlab = std::max(maxlab, data->frags13[(size_t)2 * nf + i]);
        h->init_max_label = maxlab; h->label_hi = maxlab;
        CK(h2d_sync(&h->sc->max_label, &maxlab, sizeof(int)));
        return 0;
    };
    int rc = body();
    if (rc) { g_err = h->err; ig_destroy(h); return rc; }
    *out = h;
    return 0;
}

extern "C" int ig_create(const ig_config* cfg, const ig_level_data* data, ig_handle** out) {
    if (!data) { g_err = "ig_create: null argument"; return -1; }
    return create_impl(cfg, data, nullptr, out);
}
// A further chain on the SAME level and device (replicas with different seeds, several per GPU): own scaffold, scratch,
// streams and graphs; the contacts, sub-fragment table and initial scaffold (the bulk of the memory: 8 B per contact)
// are shared with `parent` and freed when the last chain using them is destroyed.  The clone starts from the level's
// initial scaffold with no parameters set.
extern "C" int ig_clone(ig_handle* parent, ig_handle** out) {
    if (!parent || !out) { g_err = "ig_clone: null argument"; return -1; }
    return create_impl(&parent->cfg, nullptr, parent, out);
}

static void replica_free(ig_handle* h);
extern "C" void ig_destroy(ig_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    replica_free(h);
    if (h->lvl && --h->lvl->refs > 0) {   // other chains still use the level's arrays
        h->init_live = nullptr; h->sub = nullptr; h->row_ptr = nullptr; h->cv = nullptr;
        h->init_prev = nullptr; h->init_next = nullptr; h->orientable = nullptr;
    } else if (h->lvl) {
        for (void* q : {(void*)h->lvl->sym_diag, (void*)h->lvl->nb_ptr, (void*)h->lvl->nb_idx, (void*)h->lvl->nb_cdf, (void*)h->lvl->nb_nnz}) if (q) cudaFree(q);
        delete h->lvl;
    }
    h->lvl = nullptr;
    void* ptrs[] = {h->lnz_rec, h->row_dirty, h->bitmap, h->cls16, h->pick_list, h->flat_cnt, h->flat_list, h->clstab, h->subx, h->rinfo, h->cyc_frags, h->live, h->part_out, h->part_c, h->table, h->table_len, h->rowidx, h->init_live, h->sub, h->coord, h->clen, h->row_ptr, h->cv,
                    h->init_prev, h->init_next, h->orientable, h->sc, h->desc, h->exz, h->exz_test, h->chunk_cnt,
                    h->rows, h->row_cnt, h->part_nz, h->part_z, h->part_i, h->part_full, h->part_zc, h->part_nc,
                    h->d_perm, h->d_hist};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (h->h_sc) cudaFreeHost(h->h_sc);
    if (h->h_small) cudaFreeHost(h->h_small);
    if (h->h_p8) cudaFreeHost(h->h_p8);
    if (h->d_p8) cudaFree(h->d_p8);
    if (h->graph_nuis) cudaGraphExecDestroy(h->graph_nuis);
    for (int i = 0; i < 6; i++) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 16; i++) if (h->evk[i]) cudaEventDestroy(h->evk[i]);
    for (int i = 0; i < 6; i++) for (int j = 0; j <= IG_MAX_CANDS; j++) if (h->graph[i][j]) cudaGraphExecDestroy(h->graph[i][j]);
    if (h->cyc_in) cudaFree(h->cyc_in);
    if (h->cyc_out) cudaFree(h->cyc_out);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->pf) cudaStreamDestroy(h->pf);
    if (h->ev_cuts) cudaEventDestroy(h->ev_cuts);
    if (h->ev_cls) cudaEventDestroy(h->ev_cls);
    if (h->ev_pre) cudaEventDestroy(h->ev_pre);
    if (h->ev_ksc) cudaEventDestroy(h->ev_ksc);
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
    h->vinter_pos = p.v_inter > 0.0f;
    h->lnz_stale = true;  // lnz_full / z_cur depend on the parameters: the next step recomputes them (step_kind)
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
    CK(h2d_sync(&h->sc->max_label, &maxlab, sizeof(int)));
    h->label_hi = maxlab;
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
    CK(h2d_sync(h->sc->valid, in12, 12 * sizeof(int)));
    return 0;
}
extern "C" int ig_bomb(ig_handle* h, const int32_t* perm) {
    if (use(h)) return -1;
    CK(cudaMemcpyAsync(h->d_perm, perm, sizeof(int) * h->nf, cudaMemcpyHostToDevice, h->stream));
    k_explode<<<(h->nf + 255) / 256, 256, 0, h->stream>>>(h->live, h->nf, h->d_perm);
    if (launch_ok(h, "explode")) return -2;
    int maxlab = h->nf;  // perm values are 0..NF-1
    h->label_hi = maxlab;
    CK(cudaMemcpyAsync(&h->sc->max_label, &maxlab, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->coords_fresh = false;
    h->incr_valid = false;
    return 0;
}

// Contig labels are only ever compared for equality, so every applied move simply takes two fresh ones (max_label += 2).
// Long before the int32 counter could wrap (~5e8 steps) the labels are re-based to 0..NC-1 through the canonical
// relabelling of ig_get_state (what the reference does after every step, CL:2715-2806); the next step then refreshes fully.
static int labels_guard(ig_handle* h) {
    h->label_hi += 2;
    if (h->label_hi < (1LL << 30)) return 0;
    std::vector<int32_t> st((size_t)IG_N_FIELDS * h->nf);
    if (int rc = ig_get_state(h, st.data())) return rc;
    return ig_set_state(h, st.data());
}

// head of step_sampler: fill_dist_single + eval_likelihood (CL:1407-1409)
static int refresh_current(ig_handle* h, cudaStream_t st, bool fork) {
    const float mbar = h->cfg.mean_sub_len_kb;
    if (fork) {  // side stream: starts after everything already queued on the main stream
        cudaEventRecord(h->ev_fork, h->stream);
        cudaStreamWaitEvent(st, h->ev_fork, 0);
    }
    k_coords<<<h->n_part_zc, IG_THREADS, 0, st>>>(h->live, h->sub, h->coord, h->clen, h->ns, h->sc, mbar, 0, h->part_zc, h->part_nc, 1, h->subx, h->row_dirty);
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

// rows + precompute + scoring kernels of one step (after k_cand_setup / k_find_cuts / k_classes were enqueued);
// in_step: event joins with the side streams and the profiling marks of enqueue_step
#define IG_MARK(i) do { if (h->profile && !h->capturing) cudaEventRecord(h->evk[i], h->stream); } while (0)
static void enqueue_scoring(ig_handle* h, int n, bool in_step) {
    const float mbar = h->cfg.mean_sub_len_kb;
    const FragRec* live = h->live;
    if (h->rows_small) {
        if (in_step) IG_MARK(3);
        k_rows_small<<<n, IG_ROW_CHUNK, 0, h->stream>>>(h->coord, h->ns, h->sc, h->n_chunks, h->rows, h->rowidx, h->row_cnt, h->ns,
                                                        h->clstab, h->row_ptr, h->rinfo);
    } else {
        k_rows_count<<<h->n_chunks, IG_ROW_CHUNK, 0, h->stream>>>(h->coord, h->ns, h->sc, h->chunk_cnt, h->n_chunks);
        if (in_step) IG_MARK(3);
        k_rows_write<<<h->n_chunks, IG_ROW_CHUNK, 0, h->stream>>>(h->coord, h->ns, h->sc, h->chunk_cnt, h->n_chunks, h->rows,
                                                                           h->rowidx, h->row_cnt, h->ns, h->clstab, h->row_ptr, h->rinfo,
                                                                           h->bitmap, h->bitmap_words, h->cls16);
    }
    if (in_step) IG_MARK(4);
    k_precompute<<<dim3(h->grid_pre, n), IG_PRE_THREADS, 0, h->stream>>>(h->coord, h->clen, live, h->sub, h->sc, h->desc, h->rows, h->ns,
                                                                    h->table, h->table_len, mbar, h->part_z, h->part_i);
    if (h->profile && !h->capturing) cudaEventRecord(h->ev[4], h->stream);
    if (in_step) { cudaStreamWaitEvent(h->stream, h->ev_cls, 0); IG_MARK(5); }
    const int gsx = h->flat ? h->grid_flat : score_grid_x(h, n);
    if (h->flat) {
        k_pick<<<dim3(gsx, n), IG_THREADS, 0, h->stream>>>(h->cv, h->coord, h->clen, h->sc, h->desc, h->rowidx, h->ns, h->row_cnt,
                                                                    h->flat_cnt, h->chunk_stride, h->part_c, h->flat_list, h->flat_stride, mbar, h->exz,
                                                                    h->clstab, h->subx, h->rinfo);
        k_eval_flat<false><<<h->grid_score, IG_THREADS, IG_SCORE_SMEM, h->stream>>>(h->sc, h->desc, h->ns, h->flat_cnt, h->chunk_stride, h->flat_list,
                                                                            h->flat_stride, h->table, h->table_len, mbar, h->exz, h->part_nz, h->clstab, h->flat_items);
    } else {
        // (streaming mode: k_score only works for the candidates k_stream leaves to it -- circular contigs, and in the
        //  reference-faithful mode the candidates whose picks would not fit the list.  Inside a step it runs BESIDE the
        //  streaming pair on the third stream -- different candidates, different partials -- and joins before k_finalize.)
        cudaStream_t ks = h->stream;
        if (h->streaming && in_step) {
            ks = h->pf;   // (k_classes ran on it: already ordered)
            cudaEventRecord(h->ev_pre, h->stream);
            cudaStreamWaitEvent(ks, h->ev_pre, 0);
        }
        if (h->streaming) {
            k_stream<<<dim3(gsx, n), IG_THREADS, h->stream_smem, h->stream>>>(h->cv, h->coord, h->clen, h->sc, h->desc, h->rowidx, h->ns, h->row_cnt,
                                                                             h->bitmap, h->bitmap_words, h->cls16, h->part_c, h->pick_list, h->pick_cap,
                                                                             mbar, h->exz, h->clstab, h->subx, h->rinfo);
            k_eval_flat<true><<<h->grid_score, IG_THREADS, IG_SCORE_SMEM, h->stream>>>(h->sc, h->desc, h->ns, nullptr, 0, h->pick_list, h->pick_cap,
                                                                               h->table, h->table_len, mbar, h->exz, h->part_nz, h->clstab, h->flat_items);
        }
        k_score<<<dim3(gsx, n), IG_THREADS, IG_SCORE_SMEM, ks>>>(h->row_ptr, h->cv, h->coord, h->clen, h->sc, h->desc, h->rows, h->rowidx,
                                                                h->ns, h->row_cnt, h->table, h->table_len, mbar, h->exz, h->part_nz,
                                                                h->part_c, h->gs_div, h->clstab, h->subx, h->rinfo, h->sparse_div);
        if (ks != h->stream) {
            cudaEventRecord(h->ev_ksc, ks);
            cudaStreamWaitEvent(h->stream, h->ev_ksc, 0);
        }
    }
    if (h->profile && !h->capturing) cudaEventRecord(h->ev[5], h->stream);
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
    k_cand_setup<<<1, 32, 0, h->stream>>>(live, h->sc, h->desc, h->cfg.max_bounds_insert, first_flip_eject, nullptr, h->streaming ? h->stream_rows_max : 0);
    k_find_cuts<<<(h->nf + 255) / 256, 256, 0, h->stream>>>(live, h->nf, h->sc, h->desc, h->clstab);
    k_classes<<<n, IG_N_OPS * 32, 0, h->stream>>>(h->sc, h->desc, h->clstab, h->rigid, h->cfg.mean_sub_len_kb);
    if (overlap) cudaStreamWaitEvent(h->stream, h->ev_coords, 0);
    enqueue_scoring(h, n, false);
    const int gsx = h->flat ? h->grid_flat : score_grid_x(h, n);
    h->n_launches += 7;
    if (overlap) cudaStreamWaitEvent(h->stream, h->ev_lnz, 0);
    k_finalize<<<n, 1024, 0, h->stream>>>(h->row_ptr, h->cv, h->coord, h->sc, h->desc, h->rows, h->rowidx, h->ns, h->row_cnt, h->table,
                                         h->table_len, mbar, h->exz, h->part_nz, h->part_c, h->flat ? h->grid_score : gsx, h->part_z, h->part_i,
                                         h->grid_pre, h->cfg.n_pix, h->cfg.compat_last_block, h->d_nuniq, h->d_nsub, 0, gsx, h->flat ? 1 : 0);
    return launch_ok(h, "score_candidates");
}

static int apply_and_post(ig_handle* h, int forced_cand, int forced_op) {
    const int nf = h->nf;
    h->label_hi += 2;
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
    if (s.list_overflow) { h->err = "pick list of the streaming scoring path overflowed (IG_PICK_CAP)"; return -4; }
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
    if (full == 2) {
        // parameters changed, scaffold bookkeeping intact (an accepted nuisance proposal): coordinates as in the incremental
        // step (k_commit_coords flags the rows of the previous move), zero terms and the full non-zero likelihood recomputed
        // under the new parameters -- the latter from the cached per-contact records (k_lnz_refresh + k_lnz_stream)
        k_commit_coords<<<std::min(h->n_part_zc, (h->ns + 255) / 256), 256, 0, h->side>>>(h->coord, h->clen, h->sc, h->rows, h->ns,
                                                                                         h->table, h->table_len, h->live, h->sub, h->subx, h->row_dirty);
        k_coords<<<h->n_part_zc, IG_THREADS, 0, h->side>>>(h->live, h->sub, h->coord, h->clen, h->ns, h->sc, mbar, 0, h->part_zc,
                                                          h->part_nc, 0, h->subx, nullptr);
        k_reduce<<<1, 256, 0, h->side>>>(h->part_zc, h->n_part_zc, &h->sc->z_cur, h->part_nc, &h->sc->nintra_cur);
        cudaEventRecord(h->ev_coords, h->side);
        if (h->profile && !h->capturing) cudaEventRecord(h->ev[2], h->side);
        k_lnz_refresh<<<h->grid_lnz, IG_THREADS, 0, h->side>>>(h->row_ptr, h->cv, h->coord, h->clen, h->ns, h->sc, mbar, 0, h->exz,
                                                               h->lnz_rec, h->row_dirty, h->dp_bits, h->part_full + h->grid_lnz);
        k_lnz_stream<<<h->grid_lnz, IG_THREADS, 0, h->side>>>(reinterpret_cast<const int4*>(h->lnz_rec), h->lnz_pairs, h->lnz_pad,
                                                              h->lvl->val_total, h->sc, 0, h->exz, h->dp_bits, h->part_full);
        if (h->profile && !h->capturing) cudaEventRecord(h->ev[3], h->side);
        k_reduce<<<1, 256, 0, h->side>>>(h->part_full, 2 * h->grid_lnz, &h->sc->lnz_full, nullptr, nullptr);
        cudaEventRecord(h->ev_lnz, h->side);
    } else if (full) {
        k_coords<<<h->n_part_zc, IG_THREADS, 0, h->side>>>(h->live, h->sub, h->coord, h->clen, h->ns, h->sc, mbar, 0, h->part_zc,
                                                          h->part_nc, 1, h->subx, h->row_dirty);
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
                                                                                         h->table, h->table_len, h->live, h->sub, h->subx, h->row_dirty);
        cudaEventRecord(h->ev_coords, h->side);
        cudaEventRecord(h->ev_lnz, h->side);
    }
    const FragRec* live = h->live;
    IG_MARK(0);
    k_cand_setup<<<1, 32, 0, h->stream>>>(live, h->sc, h->desc, h->cfg.max_bounds_insert, 1, cycle ? h->cyc_in : nullptr, h->streaming ? h->stream_rows_max : 0);
    IG_MARK(1);
    k_find_cuts<<<(h->nf + 255) / 256, 256, 0, h->stream>>>(live, h->nf, h->sc, h->desc, h->clstab);
    cudaEventRecord(h->ev_cuts, h->stream);
    cudaStreamWaitEvent(h->pf, h->ev_cuts, 0);
    k_classes<<<n, IG_N_OPS * 32, 0, h->pf>>>(h->sc, h->desc, h->clstab, h->rigid, h->cfg.mean_sub_len_kb);   // beside the row list; k_score needs it
    cudaEventRecord(h->ev_cls, h->pf);
    cudaStreamWaitEvent(h->stream, h->ev_coords, 0);
    IG_MARK(2);
    enqueue_scoring(h, n, true);
    const int gsx = h->flat ? h->grid_flat : score_grid_x(h, n);
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
#define IG_LAUNCHES_LITE 17
// 0: incremental step; 1: coordinates and likelihood sums from scratch (scaffold replaced from outside, periodic refresh);
// 2: likelihood sums only (parameters changed; needs the per-contact records and v_inter > 0, which the records assume)
static int step_kind(const ig_handle* h) {
    if (!h->incr_valid || (h->refresh_every > 0 && h->steps_since_full >= h->refresh_every)) return 1;
    if (h->lnz_stale) return (h->lnz_cache && h->vinter_pos) ? 2 : 1;
    return 0;
}
static int step_launches(const ig_handle* h, int kind) {
    return (kind == 2 ? IG_LAUNCHES_LITE : kind ? IG_LAUNCHES_FULL : IG_LAUNCHES_INCR) - (h->rigid ? 1 : 0) + (h->prefetch ? 1 : 0) - (h->rows_small ? 1 : 0) + (h->flat ? 1 : 0) + (h->streaming ? 2 : 0);
}

static int get_graph(ig_handle* h, int full, cudaGraphExec_t* out, int cycle, int n) {
    cudaGraphExec_t& ge = h->graph[full + 3 * cycle][n];
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
    if (int rc = labels_guard(h)) return rc;
    int* hs = h->h_small;
    hs[0] = n_cands; hs[1] = id_frag;
    for (int i = 0; i < IG_MAX_CANDS; i++) hs[2 + i] = i < n_cands ? cands[i] : 0;
    const int full = step_kind(h);
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
    h->n_launches += step_launches(h, full);
    h->steps_since_full = full ? 1 : h->steps_since_full + 1;
    h->n_full += full ? 1 : 0;
    h->incr_valid = true; h->lnz_stale = false;
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
        for (int i = 3; i < 6; i++) for (int j = 0; j <= IG_MAX_CANDS; j++) if (h->graph[i][j]) { cudaGraphExecDestroy(h->graph[i][j]); h->graph[i][j] = nullptr; }  // pointers are baked in
        if (dev_alloc(h, &h->cyc_in, (size_t)n_steps * (2 + IG_MAX_CANDS)) || dev_alloc(h, &h->cyc_out, (size_t)n_steps) ||
            dev_alloc(h, &h->cyc_frags, (size_t)n_steps)) return -2;
        h->cyc_cap = n_steps;
    }
    return 0;
}

// replay n_steps steps from the plan in h->cyc_in (already on the device, or being written by an earlier kernel of
// the stream); grid_n(t) = number of candidate slots the step's grid is built for
// one step of the plan in h->cyc_in enqueued on the handle's stream (CUDA-graph replay); grid_n = number of candidate
// slots the step's grid is built for
static int enqueue_plan_step(ig_handle* h, int grid_n) {
    if (int rc = labels_guard(h)) return rc;
    const int full = step_kind(h);
    cudaGraphExec_t ge = nullptr;
    if (h->use_graph) get_graph(h, full, &ge, 1, grid_n);
    if (ge) { CK(cudaGraphLaunch(ge, h->stream)); }
    else if (enqueue_step(h, full, grid_n, 1)) return -2;
    h->steps_since_full = full ? 1 : h->steps_since_full + 1;
    h->n_full += full ? 1 : 0;
    h->incr_valid = true; h->lnz_stale = false;
    h->n_launches += step_launches(h, full);
    return 0;
}
static int begin_plan(ig_handle* h) {
    CK(cudaMemsetAsync(&h->sc->step_idx, 0, sizeof(int), h->stream));
    cudaEventRecord(h->ev[0], h->stream);
    return 0;
}
// wait for the n_steps enqueued steps and convert their device records
static int collect_plan(ig_handle* h, int n_steps, ig_cycle_step* out) {
    cudaEventRecord(h->ev[1], h->stream);
    std::vector<CycleOut> res(n_steps);
    std::vector<int> plan((size_t)n_steps * (2 + IG_MAX_CANDS));
    int overflow = 0;
    CK(cudaMemcpyAsync(res.data(), h->cyc_out, sizeof(CycleOut) * n_steps, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(plan.data(), h->cyc_in, plan.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(&overflow, &h->sc->list_overflow, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->coords_fresh = false; h->coords_ever = true;
    h->pending_steps = 0;
    if (overflow) { h->err = "pick list of the streaming scoring path overflowed (IG_PICK_CAP)"; return -4; }
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
template <class GridN>
static int run_plan(ig_handle* h, int n_steps, GridN grid_n, ig_cycle_step* out) {
    if (begin_plan(h)) return -2;
    for (int t = 0; t < n_steps; t++) if (int rc = enqueue_plan_step(h, grid_n(t))) return rc;
    return collect_plan(h, n_steps, out);
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
    LevelBlock* L = h->lvl;
    CK(cudaDeviceSynchronize());   // chains sharing the level may still be drawing from the old arrays
    for (void* q : {(void*)L->nb_ptr, (void*)L->nb_idx, (void*)L->nb_cdf, (void*)L->nb_nnz}) if (q) cudaFree(q);
    L->nb_ptr = nullptr; L->nb_idx = nullptr; L->nb_cdf = nullptr; L->nb_nnz = nullptr;
    if (dev_alloc(h, &L->nb_ptr, (size_t)h->nf + 1) || dev_alloc(h, &L->nb_idx, n + 1) || dev_alloc(h, &L->nb_cdf, n + 1) ||
        dev_alloc(h, &L->nb_nnz, (size_t)h->nf)) return -2;
    CK(h2d_sync(L->nb_ptr, ptr, sizeof(long long) * ((size_t)h->nf + 1)));
    if (n) {
        CK(h2d_sync(L->nb_idx, idx, sizeof(int) * n));
        CK(h2d_sync(L->nb_cdf, cdf, sizeof(double) * n));
    }
    CK(h2d_sync(L->nb_nnz, n_nonzero, sizeof(int) * (size_t)h->nf));
    return 0;
}

// One sweep of step_sampler over `frags` with the neighbour draws made ON THE DEVICE (Philox4x32-10 keyed by
// (seed, cycle, step, draw)): the host uploads the visiting order, one kernel draws every step's candidates,
// then the steps replay without host synchronisation.  Needs ig_set_neighbour_weights.
static int prepare_device_plan(ig_handle* h, int32_t n_steps, const int32_t* frags, int32_t n_neighbours, uint64_t seed, uint32_t cycle) {
    if (!h->params_set) { h->err = "ig_run_cycle_device: parameters not set"; return -1; }
    if (!h->lvl->nb_ptr) { h->err = "ig_run_cycle_device: neighbour weights not set (ig_set_neighbour_weights)"; return -1; }
    if (n_neighbours <= 0 || n_neighbours > IG_MAX_CANDS) { h->err = "ig_run_cycle_device: n_neighbours out of range"; return -1; }
    if (h->nf < 2) { h->err = "ig_run_cycle_device: needs at least two fragments"; return -1; }
    if (h->pending_steps) { h->err = "ig_run_cycle_device: an asynchronous cycle is still pending (ig_cycle_wait)"; return -1; }
    for (int t = 0; t < n_steps; t++) if (frags[t] < 0 || frags[t] >= h->nf) { h->err = "ig_run_cycle_device: fragment out of range"; return -1; }
    if (cycle_buffers(h, n_steps)) return -2;
    CK(cudaMemcpyAsync(h->cyc_frags, frags, sizeof(int) * (size_t)n_steps, cudaMemcpyHostToDevice, h->stream));
    k_draw_plan<<<(n_steps + 127) / 128, 128, 0, h->stream>>>(h->cyc_in, h->cyc_frags, n_steps, n_neighbours, h->nf, h->lvl->nb_ptr, h->lvl->nb_idx,
                                                            h->lvl->nb_cdf, h->lvl->nb_nnz, (unsigned)(seed & 0xffffffffu), (unsigned)(seed >> 32), cycle);
    if (launch_ok(h, "draw_plan")) return -2;
    h->n_launches += 1;
    return 0;
}
extern "C" int ig_run_cycle_device(ig_handle* h, int32_t n_steps, const int32_t* frags, int32_t n_neighbours, uint64_t seed,
                                   uint32_t cycle, ig_cycle_step* out) {
    if (use(h)) return -1;
    if (n_steps <= 0) return 0;
    if (int rc = prepare_device_plan(h, n_steps, frags, n_neighbours, seed, cycle)) return rc;
    return run_plan(h, n_steps, [&](int) { return (int)n_neighbours; }, out);
}

// Several chains of ONE device advanced together (replica chains, ig_clone): the same as one ig_run_cycle_device call per
// chain, but the steps are enqueued round-robin on the chains' own streams before anything is waited for, so the GPU runs
// them side by side -- a yeast-scale chain keeps a B200 under 25 % busy, eight of them fill it.  frags = [n_chains][n_steps]
// (each chain's visiting order), seeds = [n_chains], out = [n_chains][n_steps].  Every chain's trajectory is bit-identical
// to what ig_run_cycle_device gives it alone (tests/test_gpu_parity.py).
extern "C" int ig_run_cycles_device_multi(ig_handle** hs, int32_t n_chains, int32_t n_steps, const int32_t* frags, int32_t n_neighbours,
                                          const uint64_t* seeds, uint32_t cycle, ig_cycle_step* out) {
    if (!hs || n_chains <= 0) { g_err = "ig_run_cycles_device_multi: no chains"; return -1; }
    if (n_steps <= 0) return 0;
    for (int c = 0; c < n_chains; c++) {
        ig_handle* h = hs[c];
        if (use(h)) return -1;
        if (int rc = prepare_device_plan(h, n_steps, frags + (size_t)c * n_steps, n_neighbours, seeds[c], cycle)) { g_err = h->err; return rc; }
        if (begin_plan(h)) return -2;
    }
    // (several steps per graph launch were tried: the launching host thread is not the bottleneck, no gain)
    for (int t = 0; t < n_steps; t++)
        for (int c = 0; c < n_chains; c++)
            if (int rc = enqueue_plan_step(hs[c], n_neighbours)) { g_err = hs[c]->err; return rc; }
    int rc_all = 0;
    for (int c = 0; c < n_chains; c++)
        if (int rc = collect_plan(hs[c], n_steps, out + (size_t)c * n_steps)) { g_err = hs[c]->err; rc_all = rc; }
    return rc_all;
}
// Asynchronous form for callers that drive the chains themselves: enqueue now, collect later.
extern "C" int ig_run_cycle_device_async(ig_handle* h, int32_t n_steps, const int32_t* frags, int32_t n_neighbours, uint64_t seed, uint32_t cycle) {
    if (use(h)) return -1;
    if (n_steps <= 0) return 0;
    if (int rc = prepare_device_plan(h, n_steps, frags, n_neighbours, seed, cycle)) return rc;
    if (begin_plan(h)) return -2;
    for (int t = 0; t < n_steps; t++) if (int rc = enqueue_plan_step(h, n_neighbours)) return rc;
    h->pending_steps = n_steps;
    return 0;
}
extern "C" int ig_cycle_wait(ig_handle* h, int32_t n_steps, ig_cycle_step* out) {
    if (use(h)) return -1;
    if (n_steps != h->pending_steps) { h->err = "ig_cycle_wait: n_steps does not match the pending cycle"; return -1; }
    if (n_steps <= 0) return 0;
    return collect_plan(h, n_steps, out);
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

// Several chains on one GPU (ig_clone): each chain's scoring grids take 1 / share of the machine, so that the kernels of
// different chains run side by side instead of queueing behind each other's full-machine grids (a yeast-scale scoring
// kernel is latency-bound: 8 chains at share 4 deliver 1.2x the aggregate of full grids on T, 1.14x on the 1 Gb level).
// Changes the grouping of the f64 partial sums (not the terms): compare chains run with the SAME share bit for bit.
extern "C" int ig_set_gpu_share(ig_handle* h, int32_t share) {
    if (use(h)) return -1;
    if (share < 1) { h->err = "ig_set_gpu_share: share must be >= 1"; return -1; }
    if (h->pending_steps) { h->err = "ig_set_gpu_share: an asynchronous cycle is still pending"; return -1; }
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 6; i++) for (int j = 0; j <= IG_MAX_CANDS; j++) if (h->graph[i][j]) { cudaGraphExecDestroy(h->graph[i][j]); h->graph[i][j] = nullptr; }
    h->grid_score = std::max(8, h->grid_score_full / share);
    h->grid_flat = std::max(1, h->grid_score / 5);
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
    k_cand_setup<<<1, 32, 0, h->stream>>>(live, h->sc, h->desc, h->cfg.max_bounds_insert, 0, nullptr, 0);
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

// everything one nuisance-likelihood evaluation enqueues (test parameters in the pinned h->h_p8): used directly and under
// graph capture.  write_coords = 0: on the coordinates of the last fill (quirk Q5)
static void enqueue_full_likelihood(ig_handle* h, int write_coords, bool records) {
    const float mbar = h->cfg.mean_sub_len_kb;
    cudaMemcpyAsync(h->d_p8, h->h_p8, sizeof(Params), cudaMemcpyHostToDevice, h->stream);
    k_set_params_dev<<<1, 1, 0, h->stream>>>(h->sc, h->d_p8, 1);
    k_exz_table<<<std::min(1024, (h->ns + 256) / 256), 256, 0, h->stream>>>(h->exz_test, h->ns + 1, h->sc, mbar, 1);
    k_coords<<<h->n_part_zc, IG_THREADS, 0, h->stream>>>(h->live, h->sub, h->coord, h->clen, h->ns, h->sc, mbar, 1,
                                                        h->part_zc, h->part_nc, write_coords, h->subx, h->row_dirty);
    k_reduce<<<1, 256, 0, h->stream>>>(h->part_zc, h->n_part_zc, &h->sc->full_out[1], h->part_nc, &h->sc->full_nintra);
    cudaEventRecord(h->ev[2], h->stream);
    if (records) {
        // records of the rows whose coordinates were rewritten since the last call (+ circular contigs), then the flat pass
        k_lnz_refresh<<<h->grid_lnz, IG_THREADS, 0, h->stream>>>(h->row_ptr, h->cv, h->coord, h->clen, h->ns, h->sc, mbar, 1, h->exz_test,
                                                                 h->lnz_rec, h->row_dirty, h->dp_bits, h->part_full + h->grid_lnz);
        k_lnz_stream<<<h->grid_lnz, IG_THREADS, 0, h->stream>>>(reinterpret_cast<const int4*>(h->lnz_rec), h->lnz_pairs, h->lnz_pad,
                                                                h->lvl->val_total, h->sc, 1, h->exz_test, h->dp_bits, h->part_full);
        cudaEventRecord(h->ev[3], h->stream);
        k_reduce<<<1, 256, 0, h->stream>>>(h->part_full, 2 * h->grid_lnz, &h->sc->full_out[0], nullptr, nullptr);
    } else {
        k_full_lnz<<<h->n_part_full, IG_THREADS, 0, h->stream>>>(h->row_ptr, h->cv, h->coord, h->clen, h->ns, h->sc, mbar, 1,
                                                                h->exz_test, h->part_full);
        cudaEventRecord(h->ev[3], h->stream);
        k_reduce<<<1, 256, 0, h->stream>>>(h->part_full, h->n_part_full, &h->sc->full_out[0], nullptr, nullptr);
    }
    // full_out[3] + full_nintra: 32 contiguous bytes of the scalar record
    cudaMemcpyAsync(&h->h_sc->full_out[0], &h->sc->full_out[0], 3 * sizeof(double) + 2 * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
}

extern "C" int ig_full_likelihood(ig_handle* h, const float p8[8], int32_t use_stale_coords, double out3[3]) {
    if (use(h)) return -1;
    Params p; memcpy(&p, p8, sizeof p);
    memcpy(h->h_p8, p8, sizeof(Params));
    const int write = (use_stale_coords && h->coords_ever) ? 0 : 1;
    if (write) { h->coords_ever = true; h->coords_fresh = true; }
    const bool records = h->lnz_cache && p.v_inter > 0.0f;
    // the nuisance step's evaluation (stale coordinates, records) replays as ONE CUDA graph: 8 launches + 2 copies otherwise
    bool graphed = false;
    if (!write && records && h->use_graph && !h->nuis_graph_failed) {
        if (!h->graph_nuis) {
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
            if (e == cudaSuccess) {
                enqueue_full_likelihood(h, 0, true);
                e = cudaStreamEndCapture(h->stream, &g);
            }
            if (e == cudaSuccess && g) e = cudaGraphInstantiate(&h->graph_nuis, g, 0);
            if (g) cudaGraphDestroy(g);
            if (e != cudaSuccess || !h->graph_nuis) { h->graph_nuis = nullptr; h->nuis_graph_failed = true; cudaGetLastError(); }
        }
        if (h->graph_nuis) { CK(cudaGraphLaunch(h->graph_nuis, h->stream)); graphed = true; }
    }
    if (!graphed) enqueue_full_likelihood(h, write, records);
    if (launch_ok(h, "full_likelihood")) return -2;
    h->n_launches += records ? 7 : 6;
    CK(cudaStreamSynchronize(h->stream));
    {   // device time of the likelihood kernels (always measured: two events per call)
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]) == cudaSuccess) { h->ms_nuis += ms; h->n_nuis++; }
        else { cudaGetLastError(); if (graphed) h->nuis_graph_failed = true; }   // (events recorded by graph nodes must stay measurable)
    }
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
    k_histogram<<<h->n_part_full, IG_THREADS, 0, h->stream>>>(h->row_ptr, h->cv, h->lvl->sym_diag, h->init_live, h->sub, n_rows, bin_kb,
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
    if (!h->lvl->sym_diag) { if (dev_alloc(h, &h->lvl->sym_diag, h->ns)) return -2; }
    CK(cudaMemcpyAsync(h->lvl->sym_diag, diag, sizeof(int) * h->ns, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
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
// device self-test of the two transcribed math routines against the library ones (tests/test_gpu_math.py):
// n samples of x log-uniform in [x_lo, x_hi] with exponent y; out[0] = number of inputs where powf_pos differs from powf
// in any bit, out[1] = max |log10_f32(x) - log10((double)x)|
__global__ void k_selftest_math(int n, float x_lo, float x_hi, float y, unsigned long long* n_bad, double* max_err) {
    const float l0 = log2f(x_lo), l1 = log2f(x_hi);
    unsigned long long bad = 0;
    double worst = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned hsh = (unsigned)i * 2654435761u; hsh ^= hsh >> 15; hsh *= 2246822519u; hsh ^= hsh >> 13;
        const float x = exp2f(l0 + (l1 - l0) * ((float)(hsh >> 8) * (1.0f / 16777216.0f)));
        if (__float_as_uint(powf_pos(x, y)) != __float_as_uint(powf(x, y))) bad++;
        worst = fmax(worst, fabs(log10_f32(x) - log10((double)x)));
    }
    if (bad) atomicAdd(n_bad, bad);
    atomicMax((unsigned long long*)max_err, (unsigned long long)__double_as_longlong(worst));   // non-negative doubles order like integers
}
extern "C" int ig_selftest_math(ig_handle* h, int32_t n, float x_lo, float x_hi, float y, double out2[2]) {
    if (use(h)) return -1;
    unsigned long long* d = nullptr;
    if (dev_alloc(h, &d, 2)) return -2;
    CK(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), h->stream));
    k_selftest_math<<<296, 256, 0, h->stream>>>(n, x_lo, x_hi, y, d, reinterpret_cast<double*>(d + 1));
    if (launch_ok(h, "selftest_math")) return -2;
    unsigned long long r[2];
    CK(cudaMemcpyAsync(r, d, sizeof r, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    cudaFree(d);
    out2[0] = (double)r[0];
    memcpy(&out2[1], &r[1], sizeof(double));
    return 0;
}

extern "C" int ig_set_profiling(ig_handle* h, int32_t on) {
    if (!h) return -1;
    h->profile = on ? 1 : 0;
    return 0;
}

// device ms spent in the full-likelihood kernel by ig_full_likelihood calls (nuisance step) and their number
extern "C" int ig_get_nuisance_stats(ig_handle* h, double out2[2], int32_t reset) {
    if (!h) return -1;
    out2[0] = h->ms_nuis; out2[1] = (double)h->n_nuis;
    if (reset) { h->ms_nuis = 0.0; h->n_nuis = 0; }
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

#include "ig_replicas.cuh"   // replica chains across GPUs: NCCL all-gather inside the library

#include "ig_pyramid.cuh"   // pyramid build: binning of a level's contact list (handle-free entry point)

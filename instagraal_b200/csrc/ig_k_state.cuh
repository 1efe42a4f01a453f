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

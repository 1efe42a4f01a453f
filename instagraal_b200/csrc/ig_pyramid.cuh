// instagraal_b200 -- pyramid build: binning of a level's sparse contact list into the next level.
// Part of ig_kernels.cu (included there; not a stand-alone translation unit).
//
// The reference does this with nested Python dictionaries over every line of the contact file, once per level
// (pyramid_sparse.py:331-397 fill_sparse_pyramid_level, :686-722 the contact part of subsample_data_set).  Here: map both
// fragment ids through old -> new, order the pair, sort the 64-bit keys (stable LSD radix sort: CUB, part of the CUDA
// toolkit), sum the counts of equal keys, and -- for the HDF5 layout, whose columns keep their order of first appearance
// inside a row (`for c in list(data.keys())`, PS:378-383) -- order the groups of a row by the input position of their
// first contact.  Handle-free: works on host arrays, owns its scratch memory for the duration of the call.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cuda/functional>
#include <cuda/std/functional>

__global__ void k_bin_keys(const int* __restrict__ fa, const int* __restrict__ fb, const int* __restrict__ old2new, int n_old,
                           long long n, unsigned long long* __restrict__ key, unsigned int* __restrict__ idx, int* __restrict__ bad) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        int a = fa[i], b = fb[i];
        if (a < 0 || b < 0 || (old2new && (a >= n_old || b >= n_old))) { *bad = 1; a = 0; b = 0; }
        if (old2new) { a = old2new[a]; b = old2new[b]; }
        const unsigned lo = (unsigned)min(a, b), hi = (unsigned)max(a, b);   // mates.sort()
        key[i] = ((unsigned long long)lo << 32) | hi;
        idx[i] = (unsigned)i;
    }
}
__global__ void k_bin_gather(const int* __restrict__ nc, const unsigned int* __restrict__ idx_sorted, long long n,
                             long long* __restrict__ val_sorted) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        val_sorted[i] = (long long)nc[idx_sorted[i]];
}
// second ordering of the groups: (row, input position of the group's first contact)
__global__ void k_bin_rowpos(const unsigned long long* __restrict__ ukey, const unsigned int* __restrict__ first, long long m,
                             unsigned long long* __restrict__ key2, unsigned int* __restrict__ gid) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        key2[i] = (ukey[i] & 0xffffffff00000000ull) | first[i];
        gid[i] = (unsigned)i;
    }
}
__global__ void k_bin_emit(const unsigned long long* __restrict__ ukey, const long long* __restrict__ usum, const unsigned int* __restrict__ order,
                           long long m, int* __restrict__ out_a, int* __restrict__ out_b, long long* __restrict__ out_n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        const long long g = order ? (long long)order[i] : i;
        const unsigned long long k = ukey[g];
        out_a[i] = (int)(k >> 32); out_b[i] = (int)(k & 0xffffffffull); out_n[i] = usum[g];
    }
}

#define BK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { g_err = std::string("ig_bin_contacts: ") + cudaGetErrorString(e_); rc = -2; goto done; } } while (0)
extern "C" int ig_bin_contacts(int32_t device, int64_t n, const int32_t* fa, const int32_t* fb, const int32_t* nc,
                               const int32_t* old2new, int32_t n_old, int32_t first_appearance_order,
                               int32_t* out_a, int32_t* out_b, int64_t* out_n, int64_t* n_out) {
    int rc = 0;
    if (n < 0 || n >= (1LL << 32) || !n_out || (n > 0 && (!fa || !fb || !nc || !out_a || !out_b || !out_n))) { g_err = "ig_bin_contacts: bad argument"; return -1; }
    *n_out = 0;
    if (n == 0) return 0;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { g_err = "ig_bin_contacts: no CUDA device available (this library has no CPU path)"; return -3; }
    if (device < 0 || device >= ndev) { g_err = "ig_bin_contacts: bad device ordinal"; return -1; }
    if (cudaSetDevice(device) != cudaSuccess) { g_err = "ig_bin_contacts: cudaSetDevice failed"; return -2; }
    int *d_fa = nullptr, *d_fb = nullptr, *d_nc = nullptr, *d_map = nullptr, *d_bad = nullptr, *d_oa = nullptr, *d_ob = nullptr;
    unsigned long long *d_key = nullptr, *d_key2 = nullptr, *d_ukey = nullptr;
    unsigned int *d_idx = nullptr, *d_idx2 = nullptr, *d_first = nullptr;
    long long *d_val = nullptr, *d_usum = nullptr, *d_on = nullptr, *d_m = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0, need = 0;
    long long m = 0;
    int bad = 0;
    const int grid = 1184, blk = 256;
    BK(cudaMalloc(&d_fa, n * 4)); BK(cudaMalloc(&d_fb, n * 4)); BK(cudaMalloc(&d_nc, n * 4)); BK(cudaMalloc(&d_bad, 4));
    BK(cudaMalloc(&d_key, n * 8)); BK(cudaMalloc(&d_key2, n * 8)); BK(cudaMalloc(&d_idx, n * 4)); BK(cudaMalloc(&d_idx2, n * 4));
    BK(cudaMalloc(&d_val, n * 8)); BK(cudaMalloc(&d_ukey, n * 8)); BK(cudaMalloc(&d_usum, n * 8)); BK(cudaMalloc(&d_first, n * 4));
    BK(cudaMalloc(&d_m, 8));
    BK(cudaMemcpy(d_fa, fa, n * 4, cudaMemcpyHostToDevice)); BK(cudaMemcpy(d_fb, fb, n * 4, cudaMemcpyHostToDevice));
    BK(cudaMemcpy(d_nc, nc, n * 4, cudaMemcpyHostToDevice)); BK(cudaMemset(d_bad, 0, 4));
    if (old2new) { BK(cudaMalloc(&d_map, (size_t)std::max(n_old, 1) * 4)); BK(cudaMemcpy(d_map, old2new, (size_t)n_old * 4, cudaMemcpyHostToDevice)); }
    k_bin_keys<<<grid, blk>>>(d_fa, d_fb, d_map, n_old, n, d_key, d_idx, d_bad);
    // scratch for the CUB calls: the largest of their requirements
    cub::DeviceRadixSort::SortPairs(nullptr, need, d_key, d_key2, d_idx, d_idx2, (int)n); tmp_bytes = need;
    cub::DeviceReduce::ReduceByKey(nullptr, need, d_key2, d_ukey, d_val, d_usum, d_m, cuda::std::plus<>{}, (int)n); tmp_bytes = std::max(tmp_bytes, need);
    cub::DeviceReduce::ReduceByKey(nullptr, need, d_key2, d_ukey, d_idx2, d_first, d_m, cuda::minimum<>{}, (int)n); tmp_bytes = std::max(tmp_bytes, need);
    BK(cudaMalloc(&d_tmp, tmp_bytes));
    need = tmp_bytes;
    BK(cub::DeviceRadixSort::SortPairs(d_tmp, need, d_key, d_key2, d_idx, d_idx2, (int)n));   // stable: equal keys keep their input order
    k_bin_gather<<<grid, blk>>>(d_nc, d_idx2, n, d_val);
    need = tmp_bytes;
    BK(cub::DeviceReduce::ReduceByKey(d_tmp, need, d_key2, d_ukey, d_val, d_usum, d_m, cuda::std::plus<>{}, (int)n));
    if (first_appearance_order) { need = tmp_bytes; BK(cub::DeviceReduce::ReduceByKey(d_tmp, need, d_key2, d_ukey, d_idx2, d_first, d_m, cuda::minimum<>{}, (int)n)); }
    BK(cudaMemcpy(&m, d_m, 8, cudaMemcpyDeviceToHost));
    BK(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
    if (bad) { g_err = "ig_bin_contacts: fragment id out of range"; rc = -1; goto done; }
    m &= 0xffffffffll;   // (ReduceByKey writes an int-sized count)
    BK(cudaMalloc(&d_oa, m * 4)); BK(cudaMalloc(&d_ob, m * 4)); BK(cudaMalloc(&d_on, m * 8));
    if (first_appearance_order) {
        // groups of a row ordered by the input position of their first contact (keys reuse the sort buffers)
        k_bin_rowpos<<<grid, blk>>>(d_ukey, d_first, m, d_key, d_idx);
        need = tmp_bytes;
        BK(cub::DeviceRadixSort::SortPairs(d_tmp, need, d_key, d_key2, d_idx, d_idx2, (int)m));
        k_bin_emit<<<grid, blk>>>(d_ukey, d_usum, d_idx2, m, d_oa, d_ob, d_on);
    } else {
        k_bin_emit<<<grid, blk>>>(d_ukey, d_usum, nullptr, m, d_oa, d_ob, d_on);
    }
    BK(cudaGetLastError());
    BK(cudaMemcpy(out_a, d_oa, m * 4, cudaMemcpyDeviceToHost)); BK(cudaMemcpy(out_b, d_ob, m * 4, cudaMemcpyDeviceToHost));
    BK(cudaMemcpy(out_n, d_on, m * 8, cudaMemcpyDeviceToHost));
    *n_out = m;
done:
    for (void* q : {(void*)d_fa, (void*)d_fb, (void*)d_nc, (void*)d_map, (void*)d_bad, (void*)d_oa, (void*)d_ob, (void*)d_key, (void*)d_key2,
                    (void*)d_ukey, (void*)d_idx, (void*)d_idx2, (void*)d_first, (void*)d_val, (void*)d_usum, (void*)d_on, (void*)d_m, d_tmp})
        if (q) cudaFree(q);
    return rc;
}
#undef BK

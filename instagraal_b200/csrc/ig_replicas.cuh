// instagraal_b200 -- replica chains across GPUs: NCCL all-gather of every chain's likelihood and live scaffold, inside the
// library (no PyTorch in the product).  Part of ig_kernels.cu (host side; included last).
//
// One process per GPU owns 1..n chains of the same level (ig_create + ig_clone).  Once per cycle (or every few hundred
// steps) ig_allgather_best packs, per local chain, a 64-byte header {likelihood, n_contigs} + the live scaffold records
// (64 B x NF) straight from device memory into one send buffer, all-gathers it over NVLink (ncclAllGather on the lead
// handle's stream; NCCL is dlopen'ed, so the library loads without it), compacts the headers on the device and copies only
// those to the host: every rank then takes the same decision (best chain = highest likelihood, lowest global index on
// ties).  The winner's scaffold stays in the receive buffer of every rank (ig_get_gathered_state).
// The reference has no counterpart (single process, single GPU); a chain is inherently sequential, so there is no data-path
// collective: replicas only.
#pragma once
#include <dlfcn.h>

struct IgNcclId { char internal[128]; };
struct IgNccl {
    void* lib;
    int (*GetUniqueId)(IgNcclId*);
    int (*CommInitRank)(void**, int, IgNcclId, int);
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
    int (*CommDestroy)(void*);
    const char* (*GetErrorString)(int);
};
static IgNccl g_nccl = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
static int nccl_load(std::string& err) {
    if (g_nccl.lib) return 0;
    const char* names[] = {getenv("IG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) if (n && (lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!lib) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return -1; }
    g_nccl.GetUniqueId = (int (*)(IgNcclId*))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, IgNcclId, int))dlsym(lib, "ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(lib, "ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(lib, "ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) { err = "libnccl lacks the expected symbols"; return -1; }
    g_nccl.lib = lib;
    return 0;
}

struct ReplicaHeader { double likelihood; int n_contigs; int n_frags; long long pad[6]; };   // 64 B
struct ig_replica_state {
    void* comm;
    int rank, n_ranks, n_local;
    size_t rec_bytes;          // header + NF scaffold records
    char *send, *recv;         // [n_local] / [n_ranks * n_local] records
    ReplicaHeader* heads;      // compact headers on the device, [n_ranks * n_local]
    ReplicaHeader* h_heads;    // pinned
    cudaEvent_t ev0, ev1, ev_chain;
};

__global__ void k_replica_header(const DevScalars* __restrict__ sc, int nf, ReplicaHeader* out) {
    ReplicaHeader r;
    r.likelihood = sc->likelihood; r.n_contigs = sc->n_heads; r.n_frags = nf;
    for (int i = 0; i < 6; i++) r.pad[i] = 0;
    *out = r;
}
__global__ void k_replica_compact(const char* __restrict__ recv, size_t rec_bytes, int n, ReplicaHeader* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = *reinterpret_cast<const ReplicaHeader*>(recv + (size_t)i * rec_bytes);
}

extern "C" int ig_nccl_unique_id(char id_out[128]) {
    std::string err;
    if (nccl_load(err)) { g_err = err; return -5; }
    IgNcclId id;
    const int rc = g_nccl.GetUniqueId(&id);
    if (rc) { g_err = std::string("ncclGetUniqueId: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); return -5; }
    memcpy(id_out, id.internal, 128);
    return 0;
}

static void replica_free(ig_handle* h) {
    ig_replica_state* r = h->rep;
    if (!r) return;
    if (r->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(r->comm);
    if (r->send) cudaFree(r->send);
    if (r->recv) cudaFree(r->recv);
    if (r->heads) cudaFree(r->heads);
    if (r->h_heads) cudaFreeHost(r->h_heads);
    if (r->ev0) cudaEventDestroy(r->ev0);
    if (r->ev1) cudaEventDestroy(r->ev1);
    if (r->ev_chain) cudaEventDestroy(r->ev_chain);
    delete r;
    h->rep = nullptr;
}

// One communicator per process / GPU, owned by `lead` (any handle of the process).  n_ranks == 1 needs neither an id
// nor NCCL (the "all-gather" is a device copy): several chains on a single GPU.
extern "C" int ig_nccl_init(ig_handle* h, int32_t rank, int32_t n_ranks, int32_t n_local_chains, const char id[128]) {
    if (use(h)) return -1;
    if (n_ranks <= 0 || rank < 0 || rank >= n_ranks || n_local_chains <= 0) { h->err = "ig_nccl_init: bad rank / size"; return -1; }
    replica_free(h);
    ig_replica_state* r = new ig_replica_state();
    memset(r, 0, sizeof *r);
    h->rep = r;
    r->rank = rank; r->n_ranks = n_ranks; r->n_local = n_local_chains;
    r->rec_bytes = sizeof(ReplicaHeader) + sizeof(FragRec) * (size_t)h->nf;
    CK(cudaMalloc((void**)&r->send, r->rec_bytes * n_local_chains));
    CK(cudaMalloc((void**)&r->recv, r->rec_bytes * n_local_chains * n_ranks));
    CK(cudaMalloc((void**)&r->heads, sizeof(ReplicaHeader) * n_local_chains * n_ranks));
    CK(cudaMallocHost((void**)&r->h_heads, sizeof(ReplicaHeader) * n_local_chains * n_ranks));
    CK(cudaEventCreate(&r->ev0)); CK(cudaEventCreate(&r->ev1));
    CK(cudaEventCreateWithFlags(&r->ev_chain, cudaEventDisableTiming));
    if (n_ranks > 1) {
        if (!id) { h->err = "ig_nccl_init: null id"; return -1; }
        if (nccl_load(h->err)) return -5;
        IgNcclId nid;
        memcpy(nid.internal, id, 128);
        const int rc = g_nccl.CommInitRank(&r->comm, n_ranks, nid, rank);
        if (rc) { h->err = std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); return -5; }
    }
    return 0;
}

// All-gather of {likelihood, n_contigs, live scaffold} of the n_local chains of every rank.  lik / n_contigs =
// [n_ranks * n_local] in rank-major order; best = index of the highest likelihood (lowest index on ties), identical on
// every rank; ms = device time of pack + all-gather + header compaction (CUDA events on the lead's stream).
extern "C" int ig_allgather_best(ig_handle* h, ig_handle** local, int32_t n_local, double* lik, int32_t* n_contigs, int32_t* best,
                                 float* ms) {
    if (use(h)) return -1;
    ig_replica_state* r = h->rep;
    if (!r) { h->err = "ig_allgather_best: ig_nccl_init not called"; return -1; }
    if (n_local != r->n_local) { h->err = "ig_allgather_best: number of local chains differs from ig_nccl_init"; return -1; }
    for (int c = 0; c < n_local; c++) if (!local[c] || local[c]->nf != h->nf || local[c]->cfg.device != h->cfg.device) { h->err = "ig_allgather_best: chains must share level and device"; return -1; }
    for (int c = 0; c < n_local; c++) {   // the pack waits for everything the chain has enqueued so far
        if (local[c] == h) continue;
        CK(cudaEventRecord(r->ev_chain, local[c]->stream));
        CK(cudaStreamWaitEvent(h->stream, r->ev_chain, 0));
    }
    CK(cudaEventRecord(r->ev0, h->stream));
    for (int c = 0; c < n_local; c++) {
        char* dst = r->send + r->rec_bytes * c;
        k_replica_header<<<1, 1, 0, h->stream>>>(local[c]->sc, h->nf, reinterpret_cast<ReplicaHeader*>(dst));
        CK(cudaMemcpyAsync(dst + sizeof(ReplicaHeader), local[c]->live, sizeof(FragRec) * (size_t)h->nf, cudaMemcpyDeviceToDevice, h->stream));
    }
    const size_t bytes = r->rec_bytes * n_local;
    if (r->n_ranks > 1) {
        const int rc = g_nccl.AllGather(r->send, r->recv, bytes, /* ncclInt8 */ 0, r->comm, h->stream);
        if (rc) { h->err = std::string("ncclAllGather: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); return -5; }
    } else {
        CK(cudaMemcpyAsync(r->recv, r->send, bytes, cudaMemcpyDeviceToDevice, h->stream));
    }
    const int n_all = r->n_ranks * n_local;
    k_replica_compact<<<(n_all + 63) / 64, 64, 0, h->stream>>>(r->recv, r->rec_bytes, n_all, r->heads);
    CK(cudaEventRecord(r->ev1, h->stream));
    CK(cudaMemcpyAsync(r->h_heads, r->heads, sizeof(ReplicaHeader) * n_all, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (launch_ok(h, "allgather_best")) return -2;
    h->n_launches += 2 * n_local + 1;
    int b = 0;
    for (int i = 0; i < n_all; i++) {
        if (lik) lik[i] = r->h_heads[i].likelihood;
        if (n_contigs) n_contigs[i] = r->h_heads[i].n_contigs;
        if (r->h_heads[i].likelihood > r->h_heads[b].likelihood) b = i;
    }
    if (best) *best = b;
    if (ms) { *ms = 0.f; cudaEventElapsedTime(ms, r->ev0, r->ev1); }
    return 0;
}

// scaffold of chain `index` (rank-major, as in ig_allgather_best) from the last all-gather, canonical contig labels
extern "C" int ig_get_gathered_state(ig_handle* h, int32_t index, int32_t* out13) {
    if (use(h)) return -1;
    ig_replica_state* r = h->rep;
    if (!r || index < 0 || index >= r->n_ranks * r->n_local) { h->err = "ig_get_gathered_state: no such chain"; return -1; }
    const int nf = h->nf;
    std::vector<FragRec> tmp(nf);
    CK(cudaMemcpyAsync(tmp.data(), r->recv + r->rec_bytes * index + sizeof(ReplicaHeader), sizeof(FragRec) * nf, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < nf; i++) {
        const int* v = reinterpret_cast<const int*>(&tmp[i].f);
        for (int k = 0; k < IG_N_FIELDS; k++) out13[(size_t)k * nf + i] = v[k];
    }
    canonical_labels(nf, out13);
    return 0;
}
extern "C" int ig_nccl_finalize(ig_handle* h) {
    if (use(h)) return -1;
    replica_free(h);
    return 0;
}

// tp.h — tensor-parallel link for the LLaMA step (north_star: W_qkv/FFN-up column-split, W_o/FFN-down row-split, sum of the partials over
// NVLink).  One process per GPU; the NCCL unique id is created by rank 0 and handed to every rank by the launcher (torch.distributed broadcast
// in bench.py / tests), then passed through the C ABI.  NCCL is dlopen()ed lazily so that single-GPU use has no NCCL dependency.
//
// Two ways to sum the [n_tok, n_embd] partials of a row-split matmul:
//   * all_reduce_sum: plain ncclAllReduce (bootstrap, fallback when peer mapping is not possible)
//   * ONE-SHOT PEER ALL-REDUCE (default): every rank exports a small exchange buffer with CUDA IPC (handles travel by ncclAllGather at load);
//     the row-split matvec writes its partial into the local buffer, then ONE kernel per all-reduce signals the peers (a flag store into each
//     peer's memory over NVLink), waits for their flags, reads all partials straight out of the peers' memory and adds them to the residual in
//     rank order (every rank performs the same additions in the same order: results are identical on all ranks).  No NCCL on the data path,
//     no separate add kernel: NVSwitch gives every GPU a direct load path to every peer, a 16 KB vector is pure latency.
#pragma once
#include "common.h"

namespace mg4 {

constexpr int kTPMaxWorld = 8;
constexpr int kTPMaxRows = 8;   // rows per pass of the per-op path

struct TPPeers {             // passed by value to the kernel
    float *partial[kTPMaxWorld][2];     // [rank][buffer]: [kTPMaxRows x n_embd] partial sums (double-buffered: all-reduce n uses buffer n & 1)
    unsigned *flags[kTPMaxWorld];       // [rank]: kTPMaxWorld words; word s = the sequence number rank s has published to this rank
    int rank, world;
};

struct TPLink {
    int rank = 0, world = 1;
    void *comm = nullptr;  // ncclComm_t
    static bool unique_id(void *out128);                       // rank 0: ncclGetUniqueId
    bool init(int rank, int world, const void *id128);         // ncclCommInitRank on the current device
    void all_reduce_sum(float *buf, size_t count, cudaStream_t s);  // in place, graph-capturable (NCCL)
    // peer path: map every rank's exchange buffer (collective: all ranks call it at load with the same n_embd)
    bool setup_peers(int n_embd, cudaStream_t s);
    bool peers_ready() const { return peers_ok; }
    float *partial_out() const;   // where the next row-split matvec must write its partial (buffer of the upcoming all-reduce)
    // x[i] = resid[i] + sum over ranks of partial_r[i], i < count; graph-capturable; advances the sequence number
    void all_reduce_resid(float *x, const float *resid, size_t count, cudaStream_t s);
    void destroy();

    TPPeers peers{};
    bool peers_ok = false;
    void *local_base = nullptr;      // this rank's exchange buffer (cudaMalloc)
    void *peer_base[kTPMaxWorld] = {};
    unsigned *seq_dev = nullptr;     // all-reduces completed so far (device: graphs replay with frozen arguments)
    unsigned long long n_issued = 0; // host mirror of the sequence (selects the buffer the next matvec writes)
};

}  // namespace mg4

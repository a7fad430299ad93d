// tp.h — tensor-parallel link for the LLaMA step (north_star: W_qkv/FFN-up column-split, W_o/FFN-down row-split,
// sum of partials over NVLink).  One process per GPU; the NCCL unique id is created by rank 0 and handed to every
// rank by the launcher (torch.distributed broadcast in bench.py / tests), then passed through the C ABI.
// NCCL is dlopen()ed lazily so that single-GPU use has no NCCL dependency.
#pragma once
#include "common.h"

namespace mg4 {

struct TPLink {
    int rank = 0, world = 1;
    void *comm = nullptr;  // ncclComm_t
    static bool unique_id(void *out128);                       // rank 0: ncclGetUniqueId
    bool init(int rank, int world, const void *id128);         // ncclCommInitRank on the current device
    void all_reduce_sum(float *buf, size_t count, cudaStream_t s);  // in place, graph-capturable
    void destroy();
};

}  // namespace mg4

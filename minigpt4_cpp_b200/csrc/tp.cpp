// tp.cpp — NCCL binding through dlopen (see tp.h).
#include "tp.h"
#include <dlfcn.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace mg4 {

namespace {
typedef int ncclResult_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef ncclResult_t (*fn_get_id)(ncclUniqueId *);
typedef ncclResult_t (*fn_init_rank)(void **, int, ncclUniqueId, int);
typedef ncclResult_t (*fn_all_reduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef ncclResult_t (*fn_all_gather)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef ncclResult_t (*fn_destroy)(void *);
typedef const char *(*fn_errstr)(ncclResult_t);
struct Api { void *h = nullptr; fn_get_id get_id; fn_init_rank init_rank; fn_all_reduce all_reduce; fn_all_gather all_gather; fn_destroy destroy; fn_errstr errstr; };
Api &api() {
    static Api a;
    if (!a.h) {
        // if the host process already loaded a libnccl (e.g. torch's bundled one) reuse it, else the system one
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { a.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (a.h) break; }
        if (!a.h) MG4_PANIC("tensor parallelism requested but libnccl.so.2 cannot be loaded: %s", dlerror());
        a.get_id = (fn_get_id)dlsym(a.h, "ncclGetUniqueId");
        a.init_rank = (fn_init_rank)dlsym(a.h, "ncclCommInitRank");
        a.all_reduce = (fn_all_reduce)dlsym(a.h, "ncclAllReduce");
        a.all_gather = (fn_all_gather)dlsym(a.h, "ncclAllGather");
        a.destroy = (fn_destroy)dlsym(a.h, "ncclCommDestroy");
        a.errstr = (fn_errstr)dlsym(a.h, "ncclGetErrorString");
        if (!a.get_id || !a.init_rank || !a.all_reduce || !a.destroy) MG4_PANIC("libnccl is missing required symbols");
    }
    return a;
}
void check(ncclResult_t r, const char *what) {
    if (r != 0) MG4_PANIC("NCCL %s failed: %s", what, api().errstr ? api().errstr(r) : "?");
}
}  // namespace

bool TPLink::unique_id(void *out128) {
    ncclUniqueId id;
    check(api().get_id(&id), "ncclGetUniqueId");
    memcpy(out128, &id, 128);
    return true;
}
bool TPLink::init(int r, int w, const void *id128) {
    rank = r; world = w;
    if (w <= 1) return true;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    check(api().init_rank(&comm, w, id, r), "ncclCommInitRank");
    return true;
}
void TPLink::all_reduce_sum(float *buf, size_t count, cudaStream_t s) {
    if (world <= 1) return;
    check(api().all_reduce(buf, buf, count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, s), "ncclAllReduce");
}
// ---- one-shot peer all-reduce -------------------------------------------------------------------------------------------------------
// Exchange buffer of a rank: [2 buffers x kTPMaxRows x n_embd floats][kTPMaxWorld flag words][sequence counter]
__global__ void __launch_bounds__(256) tp_allreduce_resid_kernel(TPPeers P, unsigned *seq_ctr, int buf, float *x, const float *resid, int count) {
    const unsigned seq = *(volatile unsigned *)seq_ctr + 1u;   // this all-reduce (the counter is advanced by the last block to finish)
    if (blockIdx.x == 0 && threadIdx.x < (unsigned)P.world) {
        // publish: this rank's partial (written by the previous kernel of the stream) is complete -> tell every rank, including this one
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(P.flags[threadIdx.x] + P.rank), "r"(seq) : "memory");
    }
    if (threadIdx.x < (unsigned)P.world) {
        unsigned v, spins = 0; const unsigned *f = P.flags[P.rank] + threadIdx.x;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if ((int)(v - seq) >= 0) break;
            if (++spins > (1u << 25)) asm volatile("trap;");   // a lost peer must end in a failed launch (an error the host reports), never in a hung GPU
            __nanosleep(64);
        }
    }
    __syncthreads();
    for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) * 4; i < count; i += (int)(gridDim.x * blockDim.x) * 4) {
        float4 acc = *(const float4 *)(resid + i);
        for (int r = 0; r < P.world; ++r) {   // rank order: the same additions in the same order on every rank
            float4 v;
            asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(P.partial[r][buf] + i) : "memory");
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *(float4 *)(x + i) = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned done = atomicAdd(seq_ctr + 1, 1u);                   // blocks finished (monotonic)
        if ((done + 1u) % gridDim.x == 0u) { __threadfence(); *(volatile unsigned *)seq_ctr = seq; }   // last block: the all-reduce is complete
    }
}

bool TPLink::setup_peers(int n_embd, cudaStream_t s) {
    if (world <= 1 || world > kTPMaxWorld || getenv("MINIGPT4_B200_TP_NCCL")) return false;
    if (!api().all_gather) return false;
    const size_t part_bytes = (size_t)kTPMaxRows * (size_t)n_embd * sizeof(float);
    const size_t bytes = 2 * part_bytes + 256;
    CUDA_CHECK(cudaMalloc(&local_base, bytes)); CUDA_CHECK(cudaMemset(local_base, 0, bytes));
    CUDA_CHECK(cudaMalloc((void **)&seq_dev, 64)); CUDA_CHECK(cudaMemset(seq_dev, 0, 64));
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, local_base) != cudaSuccess) { cudaGetLastError(); MG4_ERR("tensor parallel: cudaIpcGetMemHandle failed, using NCCL all-reduce"); return false; }
    // handles travel through the communicator that already exists: all-gather of 64-byte handles (as bytes)
    unsigned char *d_in = nullptr, *d_out = nullptr;
    CUDA_CHECK(cudaMalloc((void **)&d_in, sizeof(mine))); CUDA_CHECK(cudaMalloc((void **)&d_out, sizeof(mine) * (size_t)world));
    CUDA_CHECK(cudaMemcpyAsync(d_in, &mine, sizeof(mine), cudaMemcpyHostToDevice, s));
    check(api().all_gather(d_in, d_out, sizeof(mine), /*ncclUint8*/ 1, comm, s), "ncclAllGather");
    std::vector<cudaIpcMemHandle_t> all((size_t)world);
    CUDA_CHECK(cudaMemcpyAsync(all.data(), d_out, sizeof(mine) * (size_t)world, cudaMemcpyDeviceToHost, s));
    CUDA_CHECK(cudaStreamSynchronize(s));
    cudaFree(d_in); cudaFree(d_out);
    bool ok = true;
    for (int r = 0; r < world; ++r) {
        if (r == rank) { peer_base[r] = local_base; continue; }
        if (cudaIpcOpenMemHandle(&peer_base[r], all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; peer_base[r] = nullptr; }
    }
    // every rank must take the same path: agree through a sum over the communicator
    float *flag = nullptr; CUDA_CHECK(cudaMalloc((void **)&flag, 4));
    const float mine_ok = ok ? 0.f : 1.f;
    CUDA_CHECK(cudaMemcpyAsync(flag, &mine_ok, 4, cudaMemcpyHostToDevice, s));
    all_reduce_sum(flag, 1, s);
    float bad = 0.f; CUDA_CHECK(cudaMemcpyAsync(&bad, flag, 4, cudaMemcpyDeviceToHost, s)); CUDA_CHECK(cudaStreamSynchronize(s));
    cudaFree(flag);
    if (bad != 0.f) { MG4_ERR("tensor parallel: peer mapping of the exchange buffers failed on %d rank(s), using NCCL all-reduce", (int)bad); return false; }
    peers.rank = rank; peers.world = world;
    for (int r = 0; r < world; ++r) {
        unsigned char *b = (unsigned char *)peer_base[r];
        peers.partial[r][0] = (float *)b; peers.partial[r][1] = (float *)(b + part_bytes);
        peers.flags[r] = (unsigned *)(b + 2 * part_bytes);
    }
    peers_ok = true;
    MG4_INFO("tensor parallel %d/%d: one-shot peer all-reduce over CUDA IPC mappings (%zu B exchange buffer per rank)", rank, world, bytes);
    return true;
}
float *TPLink::partial_out() const { return peers.partial[rank][(n_issued + 1) & 1]; }
void TPLink::all_reduce_resid(float *x, const float *resid, size_t count, cudaStream_t s) {
    ++n_issued;
    const int blocks = (int)std::min<size_t>(16, (count / 4 + 255) / 256);
    tp_allreduce_resid_kernel<<<blocks, 256, 0, s>>>(peers, seq_dev, (int)(n_issued & 1), x, resid, (int)count);
    CUDA_CHECK(cudaGetLastError());
}
void TPLink::destroy() {
    for (int r = 0; r < world; ++r) if (r != rank && peer_base[r]) cudaIpcCloseMemHandle(peer_base[r]);
    if (local_base) cudaFree(local_base);
    if (seq_dev) cudaFree(seq_dev);
    if (comm) { api().destroy(comm); comm = nullptr; }
}

}  // namespace mg4

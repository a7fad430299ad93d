// tp.cpp — NCCL binding through dlopen (see tp.h).
#include "tp.h"
#include <dlfcn.h>
#include <string.h>

namespace mg4 {

namespace {
typedef int ncclResult_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef ncclResult_t (*fn_get_id)(ncclUniqueId *);
typedef ncclResult_t (*fn_init_rank)(void **, int, ncclUniqueId, int);
typedef ncclResult_t (*fn_all_reduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef ncclResult_t (*fn_destroy)(void *);
typedef const char *(*fn_errstr)(ncclResult_t);
struct Api { void *h = nullptr; fn_get_id get_id; fn_init_rank init_rank; fn_all_reduce all_reduce; fn_destroy destroy; fn_errstr errstr; };
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
void TPLink::destroy() {
    if (comm) { api().destroy(comm); comm = nullptr; }
}

}  // namespace mg4

// formats.cpp — see formats.h.
#include "formats.h"
#include <chrono>
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace mg4 {

int g_verbosity = 1;
thread_local int g_load_depth = 0;

void fail(const char *fmt, ...) {
    LoadFailure f;
    va_list ap; va_start(ap, fmt); vsnprintf(f.msg, sizeof f.msg, fmt, ap); va_end(ap);
    if (g_load_depth > 0) throw f;
    fprintf(stderr, "[minigpt4-b200][fatal] %s\n", f.msg); fflush(stderr);
    abort();
}
void fail_cuda(cudaError_t e, const char *file, int line) {
    if (g_load_depth > 0) cudaGetLastError();   // (an allocation failure is not sticky: the device stays usable for the caller's next attempt)
    fail("CUDA error %s at %s:%d: %s", cudaGetErrorName(e), file, line, cudaGetErrorString(e));
}

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

static const char *kErrorNames[ErrCount] = {
    "None", "LoadModelFileHeader", "LoadModelFileVersion", "LoadModelMiniGPT4DataType", "LoadLanguageModel", "OpenImage",
    "ImageSize", "MmapSupport", "FailedToAddString", "LLamaProjectionEmbeddingInvalidSize", "FailedToAddEmbedding",
    "EosToken", "Eos", "ImageNot224_244_3", "ImageNotF32", "ImageChannelsExpectedRGB", "ImageFormatExpectedU8",
    "PathDoesNotExist", "DumpModelFileOpen", "OpenCVNotLinked"};

const char *error_name(int code) { return (code >= 0 && code < ErrCount) ? kErrorNames[code] : ""; }

// MiniGPT4DataType -> ggml type (reference minigpt4.cpp:555-739): F16,F32,I32,L64(unsupported),Q4_0..Q8_K
int container_dtype_to_gg(int dt) {
    static const int map[16] = {GG_F16, GG_F32, GG_I32, -1, GG_Q4_0, GG_Q4_1, GG_Q5_0, GG_Q5_1, GG_Q8_0, GG_Q8_1,
                                GG_Q2_K, GG_Q3_K, GG_Q4_K, GG_Q5_K, GG_Q6_K, GG_Q8_K};
    return (dt >= 0 && dt < 16) ? map[dt] : -1;
}
int gg_to_container_dtype(int gg) {
    for (int dt = 0; dt < 16; ++dt) if (container_dtype_to_gg(dt) == gg) return dt;
    return -1;
}
size_t gg_block_elems(int gg) {
    switch (gg) {
        case GG_F32: case GG_F16: case GG_I32: return 1;
        case GG_Q4_0: case GG_Q4_1: case GG_Q5_0: case GG_Q5_1: case GG_Q8_0: case GG_Q8_1: return 32;
        default: return 256;
    }
}
size_t gg_block_bytes(int gg) {
    switch (gg) {
        case GG_F32: case GG_I32: return 4;
        case GG_F16: return 2;
        case GG_Q4_0: return 18; case GG_Q4_1: return 20; case GG_Q5_0: return 22; case GG_Q5_1: return 24;
        case GG_Q8_0: return 34; case GG_Q8_1: return 40;
        case GG_Q2_K: return 84; case GG_Q3_K: return 110; case GG_Q4_K: return 144; case GG_Q5_K: return 176;
        case GG_Q6_K: return 210; case GG_Q8_K: return 292;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
MappedFile::~MappedFile() {
    if (base_) munmap(base_, size_);
    if (fd_ >= 0) close(fd_);
}
bool MappedFile::open(const std::string &path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) return false;
    struct stat st;
    if (fstat(fd_, &st) != 0) return false;
    size_ = (size_t)st.st_size;
    void *p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (p == MAP_FAILED) { base_ = nullptr; return false; }
    base_ = (uint8_t *)p;
    madvise(base_, size_, MADV_SEQUENTIAL);
    return true;
}

// bounds-checked little-endian cursor
struct Cursor {
    const uint8_t *b; size_t n; size_t pos = 0; bool ok = true;
    bool need(size_t k) { if (pos > n || k > n - pos) { ok = false; return false; } return true; }   // (no pos + k: a damaged size must not wrap)
    int32_t s4() { int32_t v = 0; if (need(4)) { memcpy(&v, b + pos, 4); pos += 4; } return v; }
    uint32_t u4() { return (uint32_t)s4(); }
    float f4() { float v = 0; if (need(4)) { memcpy(&v, b + pos, 4); pos += 4; } return v; }
    std::string str(size_t len) { std::string s; if (need(len)) { s.assign((const char *)b + pos, len); pos += len; } return s; }
    std::string lstr() { int32_t len = s4(); if (len < 0) { ok = false; return {}; } return str((size_t)len); }
};

bool json_find_int(const std::string &json, const std::string &object, const std::string &key, long *out) {
    // locate "object": { ... } at top level, then "key": <int> inside it (brace-matched, string-aware)
    size_t p = json.find("\"" + object + "\"");
    if (p == std::string::npos) return false;
    p = json.find('{', p);
    if (p == std::string::npos) return false;
    int depth = 0; size_t end = p; bool in_str = false;
    for (size_t i = p; i < json.size(); ++i) {
        char c = json[i];
        if (in_str) { if (c == '\\') ++i; else if (c == '"') in_str = false; continue; }
        if (c == '"') in_str = true;
        else if (c == '{') ++depth;
        else if (c == '}') { if (--depth == 0) { end = i; break; } }
    }
    const std::string body = json.substr(p, end - p + 1);
    size_t k = body.find("\"" + key + "\"");
    if (k == std::string::npos) return false;
    k = body.find(':', k);
    if (k == std::string::npos) return false;
    char *e = nullptr;
    long v = strtol(body.c_str() + k + 1, &e, 10);
    if (e == body.c_str() + k + 1) return false;
    *out = v;
    return true;
}

Error VisionFile::load(const std::string &path) {
    if (!file.open(path)) return ErrMmapSupport;
    Cursor c{file.data(), file.size()};
    if (c.str(4) != "ggml") { MG4_ERR("unexpected file header"); return ErrLoadModelFileHeader; }
    if (c.s4() == 0 || !c.ok) { MG4_ERR("unexpected file version"); return ErrLoadModelFileVersion; }
    file_dtype = c.s4();
    if (container_dtype_to_gg(file_dtype) < 0) return ErrLoadModelMiniGPT4DataType;
    config_json = c.lstr();
    while (c.ok && c.pos < c.n) {
        std::string mname = c.lstr();
        int32_t nt = c.s4();
        if (!c.ok || nt < 0 || (size_t)nt > (c.n - c.pos) / 12) return ErrLoadModelFileHeader;   // (a tensor record is at least 12 bytes: a wild count must not become an allocation)
        std::vector<HostTensor> metas((size_t)nt);
        for (auto &t : metas) {
            t.name = c.lstr();
            t.n_dims = c.s4();
            if (!c.ok || t.n_dims < 0 || t.n_dims > 4) return ErrLoadModelFileHeader;
            for (int d = 0; d < t.n_dims; ++d) { t.ne[d] = c.s4(); if (t.ne[d] <= 0) return ErrLoadModelFileHeader; }   // (a negative extent would wrap nbytes)
            t.gg = container_dtype_to_gg(c.s4());
            if (t.gg < 0) return ErrLoadModelMiniGPT4DataType;
            if (t.n_dims > 0 && t.ne[0] % (int64_t)gg_block_elems(t.gg)) return ErrLoadModelFileHeader;
        }
        auto &dst = models[mname];
        model_order.push_back(mname);
        for (auto &t : metas) {
            if (c.pos & 4095) c.pos = (c.pos + 4096) & ~(size_t)4095;  // blobs are page aligned
            t.nbytes = (size_t)t.nelements() / gg_block_elems(t.gg) * gg_block_bytes(t.gg);
            if (!c.need(t.nbytes)) return ErrLoadModelFileHeader;
            t.data = c.b + c.pos;
            c.pos += t.nbytes;
            dst[t.name] = t;
        }
        MG4_INFO("Model name: %s (%d tensors)", mname.c_str(), nt);
    }
    return c.ok ? ErrNone : ErrLoadModelFileHeader;
}
const HostTensor *VisionFile::find(const std::string &model, const std::string &tensor) const {
    auto m = models.find(model);
    if (m == models.end()) return nullptr;
    auto t = m->second.find(tensor);
    return t == m->second.end() ? nullptr : &t->second;
}
const HostTensor &VisionFile::get(const std::string &model, const std::string &tensor) const {
    const HostTensor *t = find(model, tensor);
    if (!t) MG4_PANIC("Couldn't find tensor %s.%s", model.c_str(), tensor.c_str());
    return *t;
}

bool LlamaFile::load(const std::string &path) {
    if (!file.open(path)) return false;
    Cursor c{file.data(), file.size()};
    if (c.u4() != 0x67676a74u) { MG4_ERR("llama file: bad magic (need ggjt)"); return false; }
    if (c.u4() != 3) { MG4_ERR("llama file: need ggjt version 3"); return false; }
    n_vocab = c.u4(); n_embd = c.u4(); n_mult = c.u4(); n_head = c.u4(); n_layer = c.u4(); n_rot = c.u4(); ftype = c.u4();
    if (!c.ok || n_vocab == 0 || n_vocab > (1u << 24) || (size_t)n_vocab > (c.n - c.pos) / 8 || n_head == 0) return false;   // (a vocabulary entry is at least 8 bytes)
    vocab.resize(n_vocab);
    for (auto &v : vocab) { uint32_t len = c.u4(); v.text = c.str(len); v.score = c.f4(); if (!c.ok) return false; }
    while (c.ok && c.pos < c.n) {
        HostTensor t;
        t.n_dims = (int)c.u4(); uint32_t name_len = c.u4(); t.gg = (int)c.u4();
        if (!c.ok || t.n_dims < 1 || t.n_dims > 2 || gg_block_bytes(t.gg) == 0) return false;
        for (int d = 0; d < t.n_dims; ++d) { t.ne[d] = c.u4(); if (t.ne[d] <= 0 || t.ne[d] > (int64_t)1 << 31) return false; }
        t.name = c.str(name_len);
        c.pos = (c.pos + 31) & ~(size_t)31;
        if (t.ne[0] % (int64_t)gg_block_elems(t.gg)) return false;
        t.nbytes = (size_t)t.nelements() / gg_block_elems(t.gg) * gg_block_bytes(t.gg);
        if (!c.need(t.nbytes)) return false;
        t.data = c.b + c.pos;
        c.pos += t.nbytes;
        tensors[t.name] = t;
    }
    return c.ok;
}
const HostTensor &LlamaFile::get(const std::string &name) const {
    auto it = tensors.find(name);
    if (it == tensors.end()) MG4_PANIC("llama file: missing tensor %s", name.c_str());
    return it->second;
}

}  // namespace mg4

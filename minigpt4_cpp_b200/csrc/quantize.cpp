// quantize.cpp — container -> container quantiser behind minigpt4_quantize_model (reference minigpt4.cpp:2817-2982).
// Offline host tool, not on the hot path.  Tensor selection rule = reference :2897-2923: F16/F32 tensors whose name
// ends in "weight", with >= 2 dims, not containing "norm"/"Norm", outside ln_vision / query_tokens / llama_proj, and
// not patch_embed.proj.weight.  Block codecs follow ggml's quantize_row_q4_0/q4_1_reference (master-31cfbb1).
// Supported targets: the 32-element block types Q4_0, Q4_1, Q5_0, Q5_1, Q8_0 (K-quants -> LoadModelMiniGPT4DataType).
#include "formats.h"
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>
#include <algorithm>

using namespace mg4;

namespace {
inline unsigned short f2h(float f) { __half h = __float2half_rn(f); unsigned short u; memcpy(&u, &h, 2); return u; }
inline float h2f(unsigned short u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

void quantize_q4_0(const float *x, unsigned char *y, size_t n) {
    for (size_t b = 0; b < n / 32; ++b, x += 32, y += 18) {
        float amax = 0.f, mx = 0.f;
        for (int j = 0; j < 32; ++j) if (amax < fabsf(x[j])) { amax = fabsf(x[j]); mx = x[j]; }
        const float d = mx / -8.f, id = d ? 1.0f / d : 0.0f;
        const unsigned short dh = f2h(d); memcpy(y, &dh, 2);
        for (int j = 0; j < 16; ++j) {
            const unsigned char a = (unsigned char)std::min(15, (int)(signed char)(x[j] * id + 8.5f));
            const unsigned char c = (unsigned char)std::min(15, (int)(signed char)(x[j + 16] * id + 8.5f));
            y[2 + j] = (unsigned char)(a | (c << 4));
        }
    }
}
void quantize_q4_1(const float *x, unsigned char *y, size_t n) {
    for (size_t b = 0; b < n / 32; ++b, x += 32, y += 20) {
        float mn = INFINITY, mx = -INFINITY;
        for (int j = 0; j < 32; ++j) { mn = std::min(mn, x[j]); mx = std::max(mx, x[j]); }
        const float d = (mx - mn) / 15.f, id = d ? 1.0f / d : 0.0f;
        const unsigned short dh = f2h(d), mh = f2h(mn); memcpy(y, &dh, 2); memcpy(y + 2, &mh, 2);
        for (int j = 0; j < 16; ++j) {
            const unsigned char a = (unsigned char)std::min(15, (int)(signed char)((x[j] - mn) * id + 0.5f));
            const unsigned char c = (unsigned char)std::min(15, (int)(signed char)((x[j + 16] - mn) * id + 0.5f));
            y[4 + j] = (unsigned char)(a | (c << 4));
        }
    }
}
// ggml quantize_row_q5_0_reference / q5_1_reference / q8_0_reference (ggml.c @ master-31cfbb1): fifth bits packed into qh, low
// half of the block in bits 0..15, high half in bits 16..31
void quantize_q5_0(const float *x, unsigned char *y, size_t n) {
    for (size_t b = 0; b < n / 32; ++b, x += 32, y += 22) {
        float amax = 0.f, mx = 0.f;
        for (int j = 0; j < 32; ++j) if (amax < fabsf(x[j])) { amax = fabsf(x[j]); mx = x[j]; }
        const float d = mx / -16.f, id = d ? 1.0f / d : 0.0f;
        const unsigned short dh = f2h(d); memcpy(y, &dh, 2);
        uint32_t qh = 0;
        for (int j = 0; j < 16; ++j) {
            const unsigned char a = (unsigned char)std::min(31, (int)(signed char)(x[j] * id + 16.5f));
            const unsigned char c = (unsigned char)std::min(31, (int)(signed char)(x[j + 16] * id + 16.5f));
            y[6 + j] = (unsigned char)((a & 0x0F) | ((c & 0x0F) << 4));
            qh |= (uint32_t)((a & 0x10) >> 4) << (j + 0);
            qh |= (uint32_t)((c & 0x10) >> 4) << (j + 16);
        }
        memcpy(y + 2, &qh, 4);
    }
}
void quantize_q5_1(const float *x, unsigned char *y, size_t n) {
    for (size_t b = 0; b < n / 32; ++b, x += 32, y += 24) {
        float mn = INFINITY, mx = -INFINITY;
        for (int j = 0; j < 32; ++j) { mn = std::min(mn, x[j]); mx = std::max(mx, x[j]); }
        const float d = (mx - mn) / 31.f, id = d ? 1.0f / d : 0.0f;
        const unsigned short dh = f2h(d), mh = f2h(mn); memcpy(y, &dh, 2); memcpy(y + 2, &mh, 2);
        uint32_t qh = 0;
        for (int j = 0; j < 16; ++j) {
            const unsigned char a = (unsigned char)((x[j] - mn) * id + 0.5f);
            const unsigned char c = (unsigned char)((x[j + 16] - mn) * id + 0.5f);
            y[8 + j] = (unsigned char)((a & 0x0F) | ((c & 0x0F) << 4));
            qh |= (uint32_t)((a & 0x10) >> 4) << (j + 0);
            qh |= (uint32_t)((c & 0x10) >> 4) << (j + 16);
        }
        memcpy(y + 4, &qh, 4);
    }
}
void quantize_q8_0(const float *x, unsigned char *y, size_t n) {
    for (size_t b = 0; b < n / 32; ++b, x += 32, y += 34) {
        float amax = 0.f;
        for (int j = 0; j < 32; ++j) amax = std::max(amax, fabsf(x[j]));
        const float d = amax / 127.f, id = d ? 1.0f / d : 0.0f;
        const unsigned short dh = f2h(d); memcpy(y, &dh, 2);
        for (int j = 0; j < 32; ++j) y[2 + j] = (unsigned char)(signed char)roundf(x[j] * id);
    }
}
typedef void (*quant_fn)(const float *, unsigned char *, size_t);
quant_fn quantizer_for(int gg) {
    switch (gg) {
        case GG_Q4_0: return quantize_q4_0; case GG_Q4_1: return quantize_q4_1;
        case GG_Q5_0: return quantize_q5_0; case GG_Q5_1: return quantize_q5_1; case GG_Q8_0: return quantize_q8_0;
    }
    return nullptr;
}
bool contains(const std::string &s, const char *v) { return s.find(v) != std::string::npos; }
bool ends_with(const std::string &s, const char *v) { const size_t m = strlen(v); return s.size() >= m && !s.compare(s.size() - m, m, v); }
void wstr(FILE *f, const std::string &s) { int32_t n = (int32_t)s.size(); fwrite(&n, 4, 1, f); fwrite(s.data(), 1, s.size(), f); }
void wint(FILE *f, int32_t v) { fwrite(&v, 4, 1, f); }
}  // namespace

extern "C" int mg4_quantize_container(const char *in_path, const char *out_path, int data_type) {
    const int out_gg = container_dtype_to_gg(data_type);
    const quant_fn qf = quantizer_for(out_gg);
    if (!qf) return ErrLoadModelMiniGPT4DataType;  // K-quants need ne[0] % 256 == 0, which no ViT-g / Q-Former matrix satisfies (1408, 768, 6144 columns only partly)
    VisionFile in;
    if (Error e = in.load(in_path)) return e;
    FILE *f = fopen(out_path, "wb");
    if (!f) return ErrDumpModelFileOpen;
    fwrite("ggml", 1, 4, f); wint(f, 1); wint(f, data_type); wstr(f, in.config_json);
    std::vector<float> tmp; std::vector<unsigned char> q;
    for (const std::string &mname : in.model_order) {
        const auto &tensors = in.models.at(mname);
        wstr(f, mname); wint(f, (int32_t)tensors.size());
        std::vector<std::pair<const HostTensor *, bool>> plan;
        for (const auto &kv : tensors) {
            const HostTensor &t = kv.second;
            const bool quant = (t.gg == GG_F16 || t.gg == GG_F32) && ends_with(t.name, "weight") && t.n_dims >= 2 && !contains(t.name, "norm") && !contains(t.name, "Norm") &&
                               mname != "ln_vision" && mname != "query_tokens" && mname != "llama_proj" && t.name != "patch_embed.proj.weight" && t.ne[0] % 32 == 0;
            plan.emplace_back(&t, quant);
            wstr(f, t.name); wint(f, t.n_dims);
            for (int d = 0; d < t.n_dims; ++d) wint(f, (int32_t)t.ne[d]);
            wint(f, quant ? data_type : gg_to_container_dtype(t.gg));
        }
        for (auto &pr : plan) {
            const HostTensor &t = *pr.first;
            long pos = ftell(f);
            if (pos & 4095) { pos = (pos + 4096) & ~4095L; fseek(f, pos, SEEK_SET); }
            if (!pr.second) { fwrite(t.data, 1, t.nbytes, f); continue; }
            const size_t n = (size_t)t.nelements();
            tmp.resize(n);
            if (t.gg == GG_F16) { const unsigned short *h = (const unsigned short *)t.data; for (size_t i = 0; i < n; ++i) tmp[i] = h2f(h[i]); }
            else memcpy(tmp.data(), t.data, n * 4);
            q.resize(n / 32 * gg_block_bytes(out_gg));
            qf(tmp.data(), q.data(), n);
            fwrite(q.data(), 1, q.size(), f);
            MG4_INFO("%s.%s | %.2f MB -> %.2f MB", mname.c_str(), t.name.c_str(), t.nbytes / 1048576.0, q.size() / 1048576.0);
        }
    }
    fclose(f);
    return ErrNone;
}

// host seam for the codec unit tests: quantise n floats (n % 32 == 0) into ggml blocks of `gg_type`; returns bytes written or -1
extern "C" long mg4_quantize_row(int gg_type, const float *x, long n, unsigned char *out) {
    const quant_fn qf = quantizer_for(gg_type);
    if (!qf || n % 32) return -1;
    qf(x, out, (size_t)n);
    return (long)((size_t)n / 32 * gg_block_bytes(gg_type));
}

// engine.h — the engine object behind the C ABI: mirrors the reference's `class MiniGPT4`
// (minigpt4.cpp:1740-2522: init, encode_image, add_tokens, add_strings, add_embedding, sample_token,
// id_to_token, reset) with a device-resident vision graph and LLaMA step instead of ggml/llama.cpp.
#pragma once
#include "llama.h"
#include "text.h"
#include "tp.h"
#include "vision.h"
#include <memory>

struct MiniGPT4Image;
struct MiniGPT4Embedding;

namespace mg4 {

struct TPConfig { bool set = false; int rank = 0, world = 1; unsigned char id[128]; };
extern TPConfig g_tp_config;

class Engine {
public:
    ~Engine();
    // path may be empty (language model only, extension entry point)
    Error init(const std::string &path, const std::string &llm_path, int verbosity, int seed, int n_ctx, int n_batch, bool numa);
    Error encode_image(const ::MiniGPT4Image *image, ::MiniGPT4Embedding *out);
    static constexpr int kEncodeLanes = 8;
    Error encode_images(const ::MiniGPT4Image *images, size_t n, ::MiniGPT4Embedding *out, float *total_ms);   // batched: concurrent lanes
    Error add_tokens(const std::vector<int32_t> &tokens);
    Error add_strings(const char *s);
    Error add_embedding(const float *rows, int n_rows);
    int32_t sample_token(const SamplingParams &p);
    const char *id_to_token(int32_t id) const;
    void reset() { n_past_ = 0; pend_ids_.clear(); pend_emb_.clear(); }
    bool flush();                     // evaluate the queued prompt rows (see add_tokens)
    static constexpr int kMaxPending = LlamaDevice::kPrefillMax;

    bool has_vision() const { return (bool)vis_; }
    VisionDevice *vision() { return vis_.get(); }
    LlamaDevice &llm() { return *llm_; }
    Tokenizer &tokenizer() { return tok_; }
    int n_past() const { return n_past_; }
    void advance(int n) { n_past_ += n; }
    int n_embd_llm() const { return llm_->dims().n_embd; }
    float last_encode_ms = 0.f;
    TPLink tp;

private:
    std::unique_ptr<VisionFile> vfile_;
    std::unique_ptr<VisionDevice> vis_;
    std::vector<std::unique_ptr<VisionDevice>> vis_lanes_;   // lanes 1.. of the batched encode (lane 0 = vis_)
    std::unique_ptr<LlamaDevice> llm_;
    Tokenizer tok_;
    std::unique_ptr<Sampler> sampler_;
    std::vector<float> logits_;
    int n_past_ = 0;
    std::vector<int32_t> pend_ids_;   // queued rows: token id, or -1 - k for row k of pend_emb_
    std::vector<float> pend_emb_;
    int pend_base_ = 0;               // position of the first queued row
    int n_batch_ = 512;
};

}  // namespace mg4

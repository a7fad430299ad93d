// engine.cpp — see engine.h.
#include "engine.h"
#include "../../include/minigpt4.h"
#include <string.h>

namespace mg4 {

TPConfig g_tp_config;

Engine::~Engine() {
    llm_.reset();
    vis_.reset();
    tp.destroy();
}

Error Engine::init(const std::string &path, const std::string &llm_path, int verbosity, int seed, int n_ctx, int n_batch, bool /*numa: CPU-only notion*/) {
    g_verbosity = verbosity;
    n_batch_ = n_batch > 0 ? n_batch : 512;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        // No CPU fallback exists: the hot path is CUDA only.
        fprintf(stderr, "[minigpt4-b200][fatal] no CUDA device available; this engine has no CPU path\n");
        abort();
    }
    if (g_tp_config.set && g_tp_config.world > 1) tp.init(g_tp_config.rank, g_tp_config.world, g_tp_config.id);
    {
        const double t0 = now_ms();
        LlamaFile lf;
        if (!lf.load(llm_path)) { MG4_ERR("failed to read language model %s", llm_path.c_str()); return ErrLoadLanguageModel; }
        tok_.init(lf.vocab);
        llm_.reset(new LlamaDevice());
        if (!llm_->load(lf, n_ctx > 0 ? n_ctx : 2048, tp.world > 1 ? &tp : nullptr)) return ErrLoadLanguageModel;
        MG4_INFO("Load language model took %.0f ms", now_ms() - t0);
    }
    if (!path.empty()) {
        const double t0 = now_ms();
        VisionFile vf;
        if (Error e = vf.load(path)) return e;
        vis_.reset(new VisionDevice());
        if (Error e = vis_->load(vf)) return e;
        if (vis_->dims().n_embd_llm != llm_->dims().n_embd)
            MG4_ERR("warning: llama_proj width %d != language model n_embd %d", vis_->dims().n_embd_llm, llm_->dims().n_embd);
        MG4_INFO("Load model from file took %.0f ms", now_ms() - t0);
    }
    sampler_.reset(new Sampler(seed));
    logits_.resize((size_t)llm_->dims().n_vocab);
    return ErrNone;
}

// input validation = reference minigpt4.cpp:2130-2138
Error Engine::encode_image(const ::MiniGPT4Image *image, ::MiniGPT4Embedding *out) {
    if (!vis_) MG4_PANIC("encode_image on a context loaded without a vision model");
    if ((long)image->width * image->height * image->channels != 224L * 224 * 3) return ErrImageNot224_244_3;
    if (image->format != MINIGPT4_IMAGE_FORMAT_F32) return ErrImageNotF32;
    const size_t n = (size_t)32 * vis_->dims().n_embd_llm;
    out->elements = n;
    out->data = new float[n];
    last_encode_ms = vis_->encode((const float *)image->data, out->data);
    MG4_INFO("Encoding image took %.3f ms on device", last_encode_ms);
    return ErrNone;
}

// add_tokens (minigpt4.cpp:2365-2382): the reference chunks by n_batch; results are batch invariant, so the device
// path uses its own chunking.  A single token goes through the captured decode graph.
Error Engine::add_tokens(const std::vector<int32_t> &tokens) {
    if (tokens.empty()) return ErrNone;
    bool ok;
    if (tokens.size() == 1) { ok = llm_->decode_step(tokens[0], n_past_); if (ok) llm_->sync(); }
    else ok = llm_->eval_tokens(tokens.data(), (int)tokens.size(), n_past_);
    if (!ok) { MG4_ERR("Failed to add string"); return ErrFailedToAddString; }
    n_past_ += (int)tokens.size();
    return ErrNone;
}
// add_strings (minigpt4.cpp:2384-2397): add_bos is ALWAYS true
Error Engine::add_strings(const char *s) { return add_tokens(tok_.encode(s ? s : "", true)); }

Error Engine::add_embedding(const float *rows, int n_rows) {
    if (!llm_->eval_embd(rows, n_rows, n_past_)) { MG4_ERR("Failed to add embedding"); return ErrFailedToAddEmbedding; }
    n_past_ += n_rows;
    return ErrNone;
}

int32_t Engine::sample_token(const SamplingParams &p) {
    if (p.temp <= 0) return llm_->argmax();  // greedy: arg-max was computed on device with the logits
    llm_->logits_to_host(logits_.data());
    return sampler_->sample(logits_.data(), (int)logits_.size(), p);
}
const char *Engine::id_to_token(int32_t id) const {  // minigpt4.cpp:2485-2497
    if (id == 2) return "</s>";
    return tok_.piece(id);
}

}  // namespace mg4

// api.cpp — the C ABI of libminigpt4.so: the reference's 18 entry points (include/minigpt4.h, mirroring reference
// minigpt4.cpp:2543-2987) plus the B200 extensions (include/minigpt4_b200.h).
#include "../../include/minigpt4_b200.h"
#include "engine.h"
#include "image.h"
#include <string.h>
#include <sys/stat.h>
#include <exception>
#include <string>

using namespace mg4;

namespace {
Engine *E(struct MiniGPT4Context *c) { return reinterpret_cast<Engine *>(c); }
bool path_exists(const char *p) { struct stat st; return p && stat(p, &st) == 0; }
// prompt constants: reference minigpt4.cpp:139-141
const char *kSystemPrompt = "Give the following image: <Img>ImageContent</Img>. You will be able to see the image once I provide it to you. Please answer my questions.###";
const char *kEosTokenSuffix = "##";
const char *kEosSuffix = "###";
}  // namespace

extern "C" {

struct MiniGPT4Context *minigpt4_model_load(const char *path, const char *llm_model, int verbosity, int seed, int n_ctx, int n_batch, bool numa) {
    g_verbosity = verbosity;
    const double t0 = now_ms();
    if (!path_exists(path)) { MG4_ERR("%s does not exist", path ? path : "(null)"); return nullptr; }
    if (!path_exists(llm_model)) { MG4_ERR("%s does not exist", llm_model ? llm_model : "(null)"); return nullptr; }
    Engine *e = new Engine();
    if (Error err = e->init(path, llm_model, verbosity, seed, n_ctx, n_batch, numa)) {
        MG4_ERR("Failed to initialize MiniGPT4: %s", error_name(err));
        delete e;
        return nullptr;
    }
    MG4_INFO("Load model from file took %.0f ms", now_ms() - t0);
    return reinterpret_cast<struct MiniGPT4Context *>(e);
}

// Image file -> 8-bit RGB (reference minigpt4.cpp:2576-2596: cv::imread(IMREAD_COLOR) + BGR2RGB, compiled only with OpenCV - its default build
// returns OpenCVNotLinked).  Decoded here without third-party code: PNG, JPEG and binary PPM / PGM (csrc/image.cpp, csrc/jpeg.cpp); anything else is OpenImage.
// Pixel buffers of MiniGPT4Image are float-array allocations whatever they hold, because minigpt4_free_image releases them as such (:2790-2798).
int minigpt4_image_load_from_file(struct MiniGPT4Context *, const char *path, struct MiniGPT4Image *image, int) {
    if (!image) return ErrOpenImage;
    try {
        RgbImage im; std::string err;
        if (!decode_image_file(path, im, err)) { MG4_ERR("%s: %s", path ? path : "(null)", err.c_str()); return ErrOpenImage; }
        const size_t bytes = im.px.size();
        float *buf = new float[(bytes + 3) / 4];
        memcpy(buf, im.px.data(), bytes);
        image->data = buf; image->width = im.w; image->height = im.h; image->channels = 3; image->format = MINIGPT4_IMAGE_FORMAT_U8;
        return ErrNone;
    } catch (const std::exception &e) {   // (an image that does not fit in host memory must not unwind through the C ABI)
        MG4_ERR("%s: %s", path ? path : "(null)", e.what());
        return ErrOpenImage;
    }
}

// 8-bit RGB of any size -> the encoder's input (reference :2598-2651): Pillow-bicubic resize to 224 x 224, / 255, CLIP mean / std, planar CHW.
// The reference describes the result as a 1 x 150528 single-channel image (it reports the shape of its concatenated planes, :2639-2641);
// minigpt4_encode_image only checks the element count (:2130), and the same description is returned here.
int minigpt4_preprocess_image(struct MiniGPT4Context *, const struct MiniGPT4Image *image, struct MiniGPT4Image *preprocessed_image, int) {
    if (!image || !preprocessed_image || !image->data) return ErrOpenImage;
    if (image->channels != 3) { MG4_ERR("Image must have 3 channels"); return ErrImageChannelsExpectedRGB; }
    if (image->format != MINIGPT4_IMAGE_FORMAT_U8) { MG4_ERR("Image must be in U8 format"); return ErrImageFormatExpectedU8; }
    if (image->width <= 0 || image->height <= 0 || image->width > 32768 || image->height > 32768) return ErrImageSize;
    float *out = nullptr;
    try {
        std::vector<uint8_t> small((size_t)224 * 224 * 3);
        resize_bicubic_u8((const uint8_t *)image->data, image->width, image->height, small.data(), 224, 224);
        out = new float[(size_t)3 * 224 * 224];
        normalize_to_chw(small.data(), 224, 224, out);
    } catch (const std::exception &e) {
        MG4_ERR("preprocess: %s", e.what());
        delete[] out;
        return ErrImageSize;
    }
    preprocessed_image->data = out; preprocessed_image->width = 1; preprocessed_image->height = 3 * 224 * 224; preprocessed_image->channels = 1;
    preprocessed_image->format = MINIGPT4_IMAGE_FORMAT_F32;
    return ErrNone;
}

int minigpt4_encode_image(struct MiniGPT4Context *ctx, struct MiniGPT4Image *image, struct MiniGPT4Embedding *embedding, size_t /*n_threads: CPU notion, ignored*/) {
    return E(ctx)->encode_image(image, embedding);
}

int minigpt4_begin_chat_image(struct MiniGPT4Context *ctx, struct MiniGPT4Embedding *image_embedding, const char *s, size_t) {
    Engine *e = E(ctx);
    if (Error err = e->add_strings("Human: <Img>")) return err;
    if (image_embedding->elements != 32u * 5120u && image_embedding->elements != 32u * 4096u) {
        MG4_ERR("LLAMA projection image embedding size not valid: %zu", image_embedding->elements);
        return ErrLLamaProjectionEmbeddingInvalidSize;
    }
    if (Error err = e->add_embedding(image_embedding->data, 32)) { MG4_ERR("Failed to add image embedding: %s", error_name(err)); return err; }
    if (Error err = e->add_strings("</Img> ")) return err;
    if (Error err = e->add_strings(s)) return err;
    if (Error err = e->add_strings("### Assistant:")) return err;
    return ErrNone;
}

int minigpt4_end_chat_image(struct MiniGPT4Context *ctx, const char **token, size_t, float temp, int32_t top_k, float top_p, float tfs_z, float typical_p,
                            int32_t /*repeat_last_n*/, float /*repeat_penalty*/, float /*alpha_presence*/, float /*alpha_frequency*/, int mirostat,
                            float mirostat_tau, float mirostat_eta, int /*penalize_nl*/) {
    Engine *e = E(ctx);
    SamplingParams sp{temp, top_k, top_p, tfs_z, typical_p, mirostat, mirostat_tau, mirostat_eta};
    const int32_t id = e->sample_token(sp);
    *token = e->id_to_token(id);
    e->add_tokens({id});  // the reference ignores this return value too (minigpt4.cpp:2715)
    return ErrNone;
}

int minigpt4_system_prompt(struct MiniGPT4Context *ctx, size_t) { return E(ctx)->add_strings(kSystemPrompt); }

int minigpt4_begin_chat(struct MiniGPT4Context *ctx, const char *s, size_t) {
    Engine *e = E(ctx);
    if (Error err = e->add_strings("Human: ")) return err;
    if (Error err = e->add_strings(s)) return err;
    if (Error err = e->add_strings("### Assistant:")) return err;
    return ErrNone;
}

int minigpt4_end_chat(struct MiniGPT4Context *ctx, const char **token, size_t n_threads, float temp, int32_t top_k, float top_p, float tfs_z, float typical_p,
                      int32_t repeat_last_n, float repeat_penalty, float alpha_presence, float alpha_frequency, int mirostat, float mirostat_tau,
                      float mirostat_eta, int penalize_nl) {
    return minigpt4_end_chat_image(ctx, token, n_threads, temp, top_k, top_p, tfs_z, typical_p, repeat_last_n, repeat_penalty, alpha_presence, alpha_frequency,
                                   mirostat, mirostat_tau, mirostat_eta, penalize_nl);
}

int minigpt4_reset_chat(struct MiniGPT4Context *ctx) { E(ctx)->reset(); return ErrNone; }

// (the reference dereferences a null argument in the next five entry points; here a null is simply "nothing to do")
int minigpt4_contains_eos_token(const char *s) { return (s && strcmp(s, kEosTokenSuffix) == 0) ? ErrEosToken : ErrNone; }
int minigpt4_is_eos(const char *s) {
    if (!s) return ErrNone;
    const size_t n = strlen(s), m = strlen(kEosSuffix);
    return (n >= m && memcmp(s + n - m, kEosSuffix, m) == 0) ? ErrEos : ErrNone;
}

int minigpt4_free(struct MiniGPT4Context *ctx) { delete E(ctx); return ErrNone; }
int minigpt4_free_image(struct MiniGPT4Image *image) {
    if (image && image->data) { delete[] (float *)image->data; image->data = nullptr; }
    return ErrNone;
}
int minigpt4_free_embedding(struct MiniGPT4Embedding *embedding) {
    if (embedding && embedding->data) { delete[] embedding->data; embedding->data = nullptr; }
    return ErrNone;
}
const char *minigpt4_error_code_to_string(int error_code) { return error_name(error_code); }
void minigpt4_set_verbosity(int verbosity) { g_verbosity = verbosity; }

// quantize.cpp
int mg4_quantize_container(const char *in_path, const char *out_path, int data_type);
int minigpt4_quantize_model(const char *in_path, const char *out_path, int data_type) {
    if (!in_path || !out_path || !path_exists(in_path)) return ErrPathDoesNotExist;
    return mg4_quantize_container(in_path, out_path, data_type);
}

// ------------------------------------------------------------------------------------------------
// extensions (include/minigpt4_b200.h)
// ------------------------------------------------------------------------------------------------
int minigpt4_b200_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
int minigpt4_b200_set_device(int device) { return cudaSetDevice(device) == cudaSuccess ? 0 : 1; }
int minigpt4_b200_tp_unique_id(void *out128) { return TPLink::unique_id(out128) ? 0 : 1; }
int minigpt4_b200_tp_configure(int rank, int world, const void *id128) {
    g_tp_config.set = world > 1; g_tp_config.rank = rank; g_tp_config.world = world;
    if (id128) memcpy(g_tp_config.id, id128, 128);
    return 0;
}
struct MiniGPT4Context *minigpt4_b200_llm_load(const char *llm_model, int n_ctx, int seed, int verbosity) {
    g_verbosity = verbosity;
    if (!path_exists(llm_model)) { MG4_ERR("%s does not exist", llm_model ? llm_model : "(null)"); return nullptr; }
    Engine *e = new Engine();
    if (Error err = e->init("", llm_model, verbosity, seed, n_ctx, 512, false)) { MG4_ERR("Failed to initialize: %s", error_name(err)); delete e; return nullptr; }
    return reinterpret_cast<struct MiniGPT4Context *>(e);
}
int minigpt4_b200_n_vocab(struct MiniGPT4Context *ctx) { return E(ctx)->llm().dims().n_vocab; }
int minigpt4_b200_n_embd(struct MiniGPT4Context *ctx) { return E(ctx)->llm().dims().n_embd; }
int minigpt4_b200_n_past(struct MiniGPT4Context *ctx) { return E(ctx)->n_past(); }
int minigpt4_b200_tokenize(struct MiniGPT4Context *ctx, const char *text, int add_bos, int32_t *out, int max_tokens) {
    std::vector<int32_t> t = E(ctx)->tokenizer().encode(text ? text : "", add_bos != 0);
    if ((int)t.size() > max_tokens) return -(int)t.size();  // llama_tokenize convention
    memcpy(out, t.data(), t.size() * sizeof(int32_t));
    return (int)t.size();
}
int minigpt4_b200_eval_tokens(struct MiniGPT4Context *ctx, const int32_t *ids, int n) { return n <= 0 ? 0 : E(ctx)->add_tokens(std::vector<int32_t>(ids, ids + n)); }
int minigpt4_b200_eval_embd(struct MiniGPT4Context *ctx, const float *rows, int n) { return E(ctx)->add_embedding(rows, n); }
int minigpt4_b200_flush(struct MiniGPT4Context *ctx) { return E(ctx)->flush() ? 0 : ErrFailedToAddString; }
int minigpt4_b200_get_logits(struct MiniGPT4Context *ctx, float *out) { E(ctx)->flush(); E(ctx)->llm().logits_to_host(out); return 0; }
int minigpt4_b200_greedy_id(struct MiniGPT4Context *ctx) { E(ctx)->flush(); return E(ctx)->llm().argmax(); }
int minigpt4_b200_get_hidden(struct MiniGPT4Context *ctx, float *out, int n_rows) { E(ctx)->flush(); E(ctx)->llm().hidden_to_host(out, n_rows); return 0; }
const char *minigpt4_b200_token_text(struct MiniGPT4Context *ctx, int32_t id) { return E(ctx)->id_to_token(id); }
int minigpt4_b200_decode_chain(struct MiniGPT4Context *ctx, int steps, int32_t *ids_out, float *ms_out) {
    Engine *e = E(ctx);
    if (!e->flush()) return ErrFailedToAddString;
    const float ms = e->llm().decode_chain(steps, e->n_past(), ids_out);
    if (ms < 0) return ErrFailedToAddString;
    e->advance(steps);
    if (ms_out) *ms_out = ms;
    return 0;
}
int minigpt4_b200_encode_images(struct MiniGPT4Context *ctx, struct MiniGPT4Images *images, struct MiniGPT4Embeddings *embeddings) {
    if (embeddings->n_embeddings < images->n_images) return ErrImageSize;
    return E(ctx)->encode_images(images->images, images->n_images, embeddings->embeddings, nullptr);
}
int minigpt4_b200_stats(struct MiniGPT4Context *ctx, struct MiniGPT4B200Stats *out) {
    Engine *e = E(ctx);
    memset(out, 0, sizeof(*out));
    const LlamaDims &d = e->llm().dims();
    out->llm_weight_bytes_per_token = (double)e->llm().weight_bytes_per_token();
    out->kernel_launches = e->llm().kernel_launches();
    if (e->has_vision()) {
        out->vision_flops_per_image = e->vision()->flops_per_image();
        out->vision_weight_bytes = (double)e->vision()->weight_bytes();
        out->kernel_launches += e->vision()->kernel_launches();
    }
    out->last_encode_ms = e->last_encode_ms;
    out->n_layer = d.n_layer; out->n_embd = d.n_embd; out->n_ff = d.n_ff; out->n_vocab = d.n_vocab; out->n_ctx = d.n_ctx;
    out->tp_rank = e->tp.rank; out->tp_world = e->tp.world; out->sm_count = e->llm().sm_count();
    out->decode_megakernel = e->llm().mega_generation();
    out->prefill_gemm = e->llm().uses_prefill_gemm() ? 1 : 0;
    return 0;
}
int minigpt4_b200_time_matvec(struct MiniGPT4Context *ctx, int kind, int reps, float *avg_ms, double *bytes_per_launch) {
    *avg_ms = E(ctx)->llm().time_matvec(kind, reps, bytes_per_launch); return 0;
}
int minigpt4_b200_tp_time_allreduce(struct MiniGPT4Context *ctx, int reps, float *us_out, int *peer_path) {
    E(ctx)->flush();
    *us_out = E(ctx)->llm().time_allreduce(reps);
    if (peer_path) *peer_path = E(ctx)->llm().tp_peer_path() ? 1 : 0;
    return 0;
}
int minigpt4_b200_mega_trace(struct MiniGPT4Context *ctx, long long *out, int max_values) { E(ctx)->flush(); return E(ctx)->llm().mega_trace(out, max_values); }
int minigpt4_b200_op_matvec(int ggml_type, int rows, int cols, const void *w_blocks, const float *x, int n, float *y) {
    LlamaDevice::test_matvec(ggml_type, rows, cols, w_blocks, x, n, y); return 0;
}
int minigpt4_b200_op_gemm_f16(int M, int T, int K, const void *w_f16, const void *x_f16, const float *bias, int epi, float *out) {
    VisionDevice::test_gemm(M, T, K, w_f16, x_f16, bias, epi, out); return 0;
}
int minigpt4_b200_op_layernorm(const float *x, int rows, int n, const float *w, const float *b, float *out) { VisionDevice::test_layernorm(x, rows, n, w, b, out); return 0; }
int minigpt4_b200_op_attention(const float *q, const float *k, const float *v, int nq, int nk, int heads, int dh, float score_div, float *out) {
    VisionDevice::test_attention(q, k, v, nq, nk, heads, dh, score_div, out); return 0;
}

int minigpt4_b200_op_dequant_f16(int ggml_type, const void *raw, long n, void *out_f16) { return VisionDevice::test_dequant(ggml_type, raw, n, out_f16); }
extern "C" long mg4_quantize_row(int gg_type, const float *x, long n, unsigned char *out);
long minigpt4_b200_host_quantize_row(int ggml_type, const float *x, long n, void *out_blocks) { return mg4_quantize_row(ggml_type, x, n, (unsigned char *)out_blocks); }

int minigpt4_b200_host_tokenize(const char *llm_model, const char *text, int add_bos, int32_t *out, int max_tokens) {
    LlamaFile f;
    if (!f.load(llm_model)) return -1000000;
    Tokenizer t; t.init(f.vocab);
    std::vector<int32_t> ids = t.encode(text ? text : "", add_bos != 0);
    if ((int)ids.size() > max_tokens) return -(int)ids.size();
    memcpy(out, ids.data(), ids.size() * sizeof(int32_t));
    return (int)ids.size();
}
int minigpt4_b200_host_sample(const float *logits, int n_vocab, int seed, float temp, int32_t top_k, float top_p, float tfs_z, float typical_p,
                              int mirostat, float mirostat_tau, float mirostat_eta, int n_draws, int32_t *out_ids) {
    Sampler s(seed);
    SamplingParams sp{temp, top_k, top_p, tfs_z, typical_p, mirostat, mirostat_tau, mirostat_eta};
    for (int i = 0; i < n_draws; ++i) out_ids[i] = s.sample(logits, n_vocab, sp);
    return 0;
}
int minigpt4_b200_host_inspect_container(const char *path, int *n_models, int *n_tensors, int *n_embd_llm) {
    VisionFile f;
    if (Error e = f.load(path)) return e;
    int nt = 0; for (auto &m : f.models) nt += (int)m.second.size();
    if (n_models) *n_models = (int)f.models.size();
    if (n_tensors) *n_tensors = nt;
    if (n_embd_llm) { const HostTensor *t = f.find("llama_proj", "weight"); *n_embd_llm = t ? (int)t->ne[1] : 0; }
    return 0;
}
int minigpt4_b200_host_inspect_ggjt(const char *path, int *n_vocab, int *n_embd, int *n_layer, int *n_tensors) {
    LlamaFile f;
    if (!f.load(path)) return ErrLoadLanguageModel;
    if (n_vocab) *n_vocab = (int)f.n_vocab;
    if (n_embd) *n_embd = (int)f.n_embd;
    if (n_layer) *n_layer = (int)f.n_layer;
    if (n_tensors) *n_tensors = (int)f.tensors.size();
    return 0;
}

}  // extern "C"

// text.h — host-side text logic of the language path: tokenizer and samplers.
// Semantics follow llama.cpp@master-31cfbb1 (llama_tokenizer, llama_sample_*), which the reference
// calls at minigpt4.cpp:2389 (llama_tokenize) and :2452-2477 (sampler chain).  Written fresh for this
// engine; the arithmetic-heavy part (logits) comes from the CUDA path, sampling stays on the host
// except greedy, which is an on-device arg-max.
#pragma once
#include "formats.h"
#include <random>
#include <unordered_map>

namespace mg4 {

class Tokenizer {
public:
    void init(const std::vector<LlamaVocabEntry> &vocab);
    // SentencePiece-style: UTF-8 characters merged greedily by best bigram score; unknown bytes -> id byte+3.
    // Empty text yields no tokens at all (not even BOS), like llama_tokenize.
    std::vector<int32_t> encode(const std::string &text, bool add_bos) const;
    const char *piece(int32_t id) const { return vocab_[(size_t)id].text.c_str(); }
    size_t size() const { return vocab_.size(); }
private:
    std::vector<LlamaVocabEntry> vocab_;
    std::unordered_map<std::string, int32_t> index_;
};

struct SamplingParams {
    float temp; int32_t top_k; float top_p; float tfs_z; float typical_p;
    int mirostat; float mirostat_tau; float mirostat_eta;
};

struct Candidate { int32_t id; float logit; float p; };

class Sampler {
public:
    explicit Sampler(int seed);
    // Mirrors MiniGPT4::sample_token (reference minigpt4.cpp:2425-2483).  Repeat/presence/frequency penalties
    // and penalize_nl are accepted by the ABI but never applied by the reference; same here.
    int32_t sample(const float *logits, int n_vocab, const SamplingParams &p);
    static int32_t greedy(const float *logits, int n_vocab);  // first arg-max
    // individual stages (exposed for unit tests)
    static void top_k(std::vector<Candidate> &c, bool &sorted, int k, size_t min_keep);
    static void softmax(std::vector<Candidate> &c, bool &sorted);
    static void top_p(std::vector<Candidate> &c, bool &sorted, float p, size_t min_keep);
    static void tail_free(std::vector<Candidate> &c, bool &sorted, float z, size_t min_keep);
    static void typical(std::vector<Candidate> &c, bool &sorted, float p, size_t min_keep);
    static void temperature(std::vector<Candidate> &c, float t);
    int32_t draw(std::vector<Candidate> &c, bool &sorted);
private:
    std::mt19937 rng_;
    float mu_v1_ = 0, mu_v2_ = 0; bool mu_v1_init_ = false, mu_v2_init_ = false;
};

}  // namespace mg4

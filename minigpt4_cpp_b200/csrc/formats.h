// formats.h — readers for the two on-disk formats on the hot path's input side.
//   * MiniGPT-4 "ggml" container: writer reference minigpt4/convert.py:56-180, reader minigpt4.cpp:1478-1596.
//   * LLaMA ggjt v3: llama.cpp@master-31cfbb1 llama_file_loader.
// Both are mapped read-only; tensors are views into the mapping and are uploaded (and repacked for
// 128-bit loads) to HBM by the device-side loaders — the mapping is dropped after upload.
#pragma once
#include "common.h"
#include <map>
#include <memory>

namespace mg4 {

class MappedFile {
public:
    ~MappedFile();
    bool open(const std::string &path);
    const uint8_t *data() const { return base_; }
    size_t size() const { return size_; }
private:
    uint8_t *base_ = nullptr; size_t size_ = 0; int fd_ = -1;
};

// minimal JSON value lookups for the embedded config (only three integers are consumed:
// Qformer.encoder_width / query_length / num_hidden_layers — reference minigpt4.cpp:2146,:2227,:2293)
bool json_find_int(const std::string &json, const std::string &object, const std::string &key, long *out);

struct VisionFile {
    MappedFile file;
    int file_dtype = 0;                       // MiniGPT4DataType of the file header
    std::string config_json;
    std::vector<std::string> model_order;     // sub-model names in file order
    std::map<std::string, std::map<std::string, HostTensor>> models;  // model -> tensor name -> view
    Error load(const std::string &path);
    const HostTensor *find(const std::string &model, const std::string &tensor) const;
    const HostTensor &get(const std::string &model, const std::string &tensor) const;  // PANIC if missing (reference :898)
};

struct LlamaVocabEntry { std::string text; float score; };

struct LlamaFile {
    MappedFile file;
    uint32_t n_vocab = 0, n_embd = 0, n_mult = 0, n_head = 0, n_layer = 0, n_rot = 0, ftype = 0;
    std::vector<LlamaVocabEntry> vocab;
    std::map<std::string, HostTensor> tensors;
    bool load(const std::string &path);  // false -> LoadLanguageModel
    const HostTensor &get(const std::string &name) const;
    uint32_t n_ff() const { return ((2 * (4 * n_embd) / 3 + n_mult - 1) / n_mult) * n_mult; }
};

}  // namespace mg4

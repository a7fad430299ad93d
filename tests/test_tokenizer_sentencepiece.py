"""External pin for the tokenizer (host logic, CPU): llama.cpp's `llama_tokenize` (reference call site minigpt4.cpp:2389) is a
restatement of sentencepiece's BPE encoder over the vocabulary that convert.py copies out of `tokenizer.model` ('▁' -> ' ', byte pieces ->
raw bytes, scores kept).  Here a small BPE model is trained with the sentencepiece library itself (LLaMA's layout: <unk>/<s>/</s>, 256 byte
pieces, byte fallback, identity normalisation), converted the same way, written into a ggjt file, and the product tokenizer (C++, through the
C ABI) and the oracle's Python restatement must reproduce sentencepiece's ids.  llama.cpp callers prepend the dummy-prefix space themselves,
so `" " + text` here corresponds to sentencepiece's add_dummy_prefix."""
from pathlib import Path

import numpy as np
import pytest

spm = pytest.importorskip("sentencepiece")
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def sp_setup(tmp_path_factory, mg):
    d = tmp_path_factory.mktemp("spm")
    lines = []
    for f in ("SURVEY.md", "DESIGN.md", "README.md", "INTEGRATION.md"):
        lines += [t for t in (ROOT / f).read_text(encoding="utf-8").split("\n") if t.strip()]
    corpus = d / "corpus.txt"
    corpus.write_text("\n".join(lines), encoding="utf-8")
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(d / "tiny"), vocab_size=1200, model_type="bpe", byte_fallback=True,
                                   normalization_rule_name="identity", remove_extra_whitespaces=False, add_dummy_prefix=True,
                                   character_coverage=0.995, unk_id=0, bos_id=1, eos_id=2, pad_id=-1, minloglevel=2)
    sp = spm.SentencePieceProcessor(model_file=str(d / "tiny.model"))
    vocab = []
    for i in range(sp.get_piece_size()):  # llama.cpp convert.py, SentencePieceVocab.sentencepiece_tokens
        p = sp.id_to_piece(i)
        if sp.is_unknown(i):
            t = " ⁇ ".encode("utf-8")
        elif sp.is_control(i):
            t = b""
        elif sp.is_byte(i):
            t = bytes([int(p[3:-1], 16)])
        else:
            t = p.replace("▁", " ").encode("utf-8")
        vocab.append((t, float(sp.get_score(i))))
    llm = str(d / "llama-spm.bin")
    mg.write_llama_ggjt(llm, mg.LlamaSpec(n_vocab=len(vocab), n_embd=128, n_head=1, n_layer=1, wtype="f16", vocab=vocab))
    return sp, vocab, llm, lines


def _samples(lines):
    rng = np.random.default_rng(5)
    out = ["Human: <Img>", "</Img> ", "### Assistant:", "what is in this picture?", "héllo wörld ✓ naïve café", "  two  spaces ", "a", "###",
           "Give the following image: <Img>ImageContent</Img>. You will be able to see the image once I provide it to you.",
           "tcgen05.mma cta_group::2 pairs two SMs", "日本語 text with CJK", "tab\tand\nnewline", "UPPER lower MiXeD 12345 67.89"]
    out += [lines[int(i)][:200] for i in rng.integers(0, len(lines), 150)]
    for _ in range(100):
        out.append(bytes(rng.integers(32, 127, size=int(rng.integers(1, 60))).tolist()).decode())
    return out


def test_product_and_oracle_tokenizers_reproduce_sentencepiece(ext, orc, sp_setup):
    sp, vocab, llm, lines = sp_setup
    tk = orc.Tokenizer(vocab)
    n = 0
    for text in _samples(lines):
        want = sp.encode(text)
        assert tk.tokenize(" " + text, False) == want, text
        assert ext.host_tokenize(llm, " " + text, False) == want, text
        assert ext.host_tokenize(llm, " " + text, True) == [1] + want
        n += len(want)
    assert n > 5000

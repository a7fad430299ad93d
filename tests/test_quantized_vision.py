"""Quantised MiniGPT-4 containers (minigpt4_quantize_model output; the README's pre-quantised downloads).

Host side (CPU tests): the block quantisers behind `minigpt4_quantize_model` (reference minigpt4.cpp:2817-2982, ggml
quantize_row_*_reference) against gguf-py; the container round trip for every supported target; and the size of the one deliberate
numerical deviation — the vision graph expands a quantised matrix to F16 at load time instead of running ggml's quantised mul_mat
(activations quantised to Q8) — measured with the oracle alone: quantised container vs its dequantised-F16 twin.
Device side (GPU tests): the load-time expansion kernel is exact against gguf-py's dequantiser, and the encode of a quantised
container matches the oracle."""
import numpy as np
import pytest

from conftest import rel_err

# ggml type id -> (gguf name, container dtype)
TYPES = {2: ("Q4_0", 4), 3: ("Q4_1", 5), 6: ("Q5_0", 6), 7: ("Q5_1", 7), 8: ("Q8_0", 8)}


def _gguf(gt):
    import gguf
    return getattr(gguf.GGMLQuantizationType, TYPES[gt][0])


@pytest.mark.parametrize("gt", sorted(TYPES))
def test_block_quantisers_match_gguf(ext, gt):
    import gguf.quants as gq
    rng = np.random.default_rng(gt)
    x = (rng.standard_normal((8, 1408)) * 0.02).astype(np.float32)
    x[0, :32] = 0.0                       # an all-zero block (d == 0 -> id = 0)
    x[1, 5] = 3.0                         # an outlier block
    got = ext.host_quantize_row(gt, x)
    ref = gq.quantize(x, _gguf(gt)).reshape(-1)
    assert got.size == ref.size
    assert np.mean(got == ref) > 0.99     # identical up to float rounding of 1/d in a handful of codes
    back = gq.dequantize(got.reshape(8, -1), _gguf(gt))
    step = {2: 1 / 8, 3: 1 / 15, 6: 1 / 16, 7: 1 / 31, 8: 1 / 127}[gt]
    for r in range(8):
        for b in range(0, 1408, 32):
            blk = x[r, b:b + 32]
            span = (blk.max() - blk.min()) if gt in (3, 7) else np.abs(blk).max()
            assert np.abs(back[r, b:b + 32] - blk).max() <= span * step * 1.05 + 1e-6   # (Q4_0 / Q5_0 clip the side opposite to the extreme value: up to one step)
    assert ext.L.minigpt4_b200_host_quantize_row(10, None, 32, None) == -1   # K-quants: not a vision-container target
    assert ext.L.minigpt4_b200_host_quantize_row(3, None, 33, None) == -1    # ragged block


def _twin_f16(orc, mg, src, dst):
    """Rewrite a (quantised) container with every quantised matrix replaced by its dequantised values rounded to F16."""
    import gguf.quants as gq
    import json, struct
    mm = np.memmap(src, dtype=np.uint8, mode="r")
    n = struct.unpack_from("<i", mm, 12)[0]
    config = json.loads(bytes(mm[16:16 + n]).decode())
    _, tensors = orc.read_minigpt4(src)
    models: dict[str, dict[str, np.ndarray]] = {}
    for key, tv in tensors.items():
        mname = next(m for m in ("visual_encoder", "ln_vision", "query_tokens", "Qformer", "llama_proj") if key.startswith(m + "."))
        tname = key[len(mname) + 1:]
        shape = list(tv.ne)[::-1]
        if tv.gtype == 0:
            arr = np.frombuffer(tv.data, np.float32).reshape(shape)
        elif tv.gtype == 1:
            arr = np.frombuffer(tv.data, np.float16).reshape(shape)
        else:
            raw = np.frombuffer(tv.data, np.uint8).reshape(int(np.prod(shape[:-1])), -1)
            arr = gq.dequantize(raw, _gguf(tv.gtype)).astype(np.float16).reshape(shape)
        models.setdefault(mname, {})[tname] = arr
    order = ["visual_encoder", "ln_vision", "query_tokens", "Qformer", "llama_proj"]
    mg.write_container(dst, config, [(m, models[m]) for m in order])


@pytest.mark.parametrize("gt", sorted(TYPES))
def test_quantize_model_every_target(lib, orc, tiny, tmp_path, gt):
    import gguf.quants as gq
    out = str(tmp_path / f"q{gt}.bin")
    lib.minigpt4_quantize_model(tiny["vision"], out, TYPES[gt][1])
    _, src = orc.read_minigpt4(tiny["vision"])
    _, dst = orc.read_minigpt4(out)
    assert set(src) == set(dst)
    qn = "Qformer.bert.encoder.layer.0.attention.self.query.weight"
    assert dst[qn].gtype == gt and src[qn].gtype == 1
    for keep in ("visual_encoder.patch_embed.proj.weight", "llama_proj.weight", "ln_vision.weight", "query_tokens.weight"):
        assert dst[keep].gtype == src[keep].gtype and bytes(dst[keep].data) == bytes(src[keep].data)
    rows, cols = src[qn].ne[1], src[qn].ne[0]
    w = np.frombuffer(src[qn].data, np.float16).reshape(rows, cols).astype(np.float32)
    wq = gq.dequantize(np.frombuffer(dst[qn].data, np.uint8).reshape(rows, -1), _gguf(gt))
    levels = {2: 8, 3: 15, 6: 16, 7: 31, 8: 127}[gt]
    assert np.abs(w - wq).max() <= (w.max() - w.min()) / levels * 1.05


@pytest.mark.parametrize("gt", sorted(TYPES))
def test_oracle_block_types_match_gguf(orc, gt):
    """The oracle's restatement of every block type a quantised container can hold: dequantiser exact against gguf-py, quantised
    mul_mat (ggml_vec_dot_q*_q8_*: Q8 activations, integer block dots) within the activation-quantisation error of the float product."""
    import gguf.quants as gq
    rng = np.random.default_rng(40 + gt)
    w = (rng.standard_normal((12, 1408)) * 0.02).astype(np.float32)
    raw = gq.quantize(w, _gguf(gt))
    ref = gq.dequantize(raw, _gguf(gt))
    assert np.array_equal(orc.dequant_rows(gt, raw, 12, 1408), ref)
    x = rng.standard_normal((5, 1408)).astype(np.float32)
    assert rel_err(orc.mul_mat(gt, raw, 12, 1408, x), x @ ref.T) < 1.5e-2


@pytest.mark.parametrize("gt", sorted(TYPES))
def test_oracle_quantised_vs_dequantised_twin(lib, orc, mg, tiny, tmp_path, gt):
    """How far ggml's quantised mul_mat (Q8 activations x quantised weights) is from 'same stored weights, F16 operands' on the
    embedding: the gap the GPU path inherits by expanding quantised matrices to F16.  Must sit well inside the 1e-2 parity bar
    (measured 4.5e-3 .. 5.3e-3 for all five types on this 2-block model: the activation quantiser dominates, not the weight width)."""
    q = str(tmp_path / "q41.bin"); twin = str(tmp_path / "q41-f16.bin")
    lib.minigpt4_quantize_model(tiny["vision"], q, TYPES[gt][1])
    _twin_f16(orc, mg, q, twin)
    img = mg.synth_image(3)
    a = orc.OracleEngine(q, None).encode_image(img)
    b = orc.OracleEngine(twin, None).encode_image(img)
    assert a.shape == b.shape == (32, 4096)
    assert rel_err(a, b) < 6e-3, rel_err(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("gt", [0] + sorted(TYPES))
def test_load_time_expansion_is_exact(ext, gt):
    import gguf.quants as gq
    rng = np.random.default_rng(100 + gt)
    x = (rng.standard_normal((6, 1408)) * 0.05).astype(np.float32)
    if gt == 0:
        got = ext.op_dequant_f16(0, x, x.size)
        assert np.array_equal(got.view(np.uint16), x.astype(np.float16).reshape(-1).view(np.uint16))
        return
    raw = gq.quantize(x, _gguf(gt))
    want = gq.dequantize(raw, _gguf(gt)).astype(np.float16).reshape(-1)
    got = ext.op_dequant_f16(gt, raw, x.size)
    assert np.array_equal(got.view(np.uint16), want.view(np.uint16))


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [5, 7])  # Q4_1, Q5_1 containers
def test_encode_quantised_container(lib, ext, orc, mg, tiny, tmp_path, dt):
    q = str(tmp_path / "q.bin"); twin = str(tmp_path / "q-f16.bin")
    lib.minigpt4_quantize_model(tiny["vision"], q, dt)
    _twin_f16(orc, mg, q, twin)
    llm = str(tmp_path / "llama-4096.bin")
    mg.write_llama_ggjt(llm, mg.LlamaSpec(n_vocab=512, n_embd=4096, n_head=32, n_layer=1, wtype="q4_1"))
    c = lib.minigpt4_model_load(q, llm, 1, 1, 64, 8, 0)
    assert c.ptr
    img = mg.synth_image(11)
    got = ext.encode_array(c, img)
    lib.minigpt4_free(c)
    same_weights = orc.OracleEngine(twin, None).encode_image(img)      # same stored weights as F16 operands: the F16-path tolerance
    assert rel_err(got, same_weights) < 2e-3, rel_err(got, same_weights)
    ggml_way = orc.OracleEngine(q, None).encode_image(img)             # ggml's quantised mul_mat (Q8 activations, integer dots)
    assert rel_err(got, ggml_way) < 1e-2, rel_err(got, ggml_way)

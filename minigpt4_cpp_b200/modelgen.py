"""Synthetic-weight generators for the two on-disk formats the hot path consumes.

* MiniGPT-4 ``"ggml"`` container — layout per reference ``minigpt4/convert.py:56-180`` (writer)
  and ``minigpt4.cpp:1478-1596`` (reader): magic, version, dtype, config JSON, then per sub-model
  a tensor table followed by 4096-aligned raw blobs.  Shapes are stored REVERSED (ggml ``ne`` order).
* LLaMA **ggjt v3** — llama.cpp@master-31cfbb1 file format: magic/version, 7 hparams,
  vocab (len, bytes, score), then tensors with 32-byte aligned data.

There is no network for real checkpoints, so weights are seeded random.  Quantised tensors are
synthesised *directly in the quantised domain* (random nibbles / high bits / scales): the
dequantised values ARE the model, so ggml's quantisation search never has to be reproduced.
"""
from __future__ import annotations

import json
import struct
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

# MiniGPT4DataType ids (include/minigpt4.h; reference minigpt4.h:30-48)
DT_F16, DT_F32, DT_I32, DT_L64 = 0, 1, 2, 3
DT_Q4_0, DT_Q4_1, DT_Q5_K, DT_Q6_K = 4, 5, 13, 14
# ggml_type ids used inside ggjt files
GG_F32, GG_F16, GG_Q4_0, GG_Q4_1, GG_Q5_K, GG_Q6_K = 0, 1, 2, 3, 13, 14
GG_Q5_0, GG_Q5_1, GG_Q8_0, GG_Q4_K = 6, 7, 8, 12  # oracle / file-format support only so far: no CUDA matvec for these on the LLaMA path yet
GG_NAMES = {"f32": GG_F32, "f16": GG_F16, "q4_0": GG_Q4_0, "q4_1": GG_Q4_1, "q5_k": GG_Q5_K, "q6_k": GG_Q6_K,
            "q5_0": GG_Q5_0, "q5_1": GG_Q5_1, "q8_0": GG_Q8_0, "q4_k": GG_Q4_K}
GG_BLOCK = {GG_F32: (1, 4), GG_F16: (1, 2), GG_Q4_0: (32, 18), GG_Q4_1: (32, 20), GG_Q5_K: (256, 176), GG_Q6_K: (256, 210),
            GG_Q5_0: (32, 22), GG_Q5_1: (32, 24), GG_Q8_0: (32, 34), GG_Q4_K: (256, 144)}


def gg_row_bytes(gtype: int, cols: int) -> int:
    per, nbytes = GG_BLOCK[gtype]
    assert cols % per == 0
    return cols // per * nbytes


# ----------------------------------------------------------------------------------------------
# quantised-domain synthesis
# ----------------------------------------------------------------------------------------------
def synth_quant(rng: np.random.Generator, gtype: int, rows: int, cols: int, sigma: float) -> np.ndarray:
    """Return uint8 [rows, row_bytes] holding blocks whose dequantised values have std ~= sigma."""
    if gtype == GG_F32:
        return (rng.standard_normal((rows, cols), dtype=np.float32) * sigma).view(np.uint8).reshape(rows, -1)
    if gtype == GG_F16:
        return (rng.standard_normal((rows, cols), dtype=np.float32) * sigma).astype(np.float16).view(np.uint8).reshape(rows, -1)
    per, nb = GG_BLOCK[gtype]
    nblk = rows * cols // per
    out = np.empty((nblk, nb), dtype=np.uint8)
    jitter = (0.75 + 0.5 * rng.random(nblk, dtype=np.float32))
    if gtype == GG_Q4_1:
        # w = d*q + m, q uniform 0..15 -> std = d*sqrt(255/12); centre with m = -7.5 d (+ small offset)
        d = (sigma / 4.61 * jitter).astype(np.float16)
        m = (-7.5 * d.astype(np.float32) * (0.9 + 0.2 * rng.random(nblk, dtype=np.float32))).astype(np.float16)
        out[:, 0:2] = d.view(np.uint8).reshape(nblk, 2)
        out[:, 2:4] = m.view(np.uint8).reshape(nblk, 2)
        out[:, 4:] = rng.integers(0, 256, size=(nblk, 16), dtype=np.uint8)
    elif gtype == GG_Q4_0:
        d = (sigma / 4.61 * jitter).astype(np.float16)
        out[:, 0:2] = d.view(np.uint8).reshape(nblk, 2)
        out[:, 2:] = rng.integers(0, 256, size=(nblk, 16), dtype=np.uint8)
    elif gtype == GG_Q5_K:
        # w = d*sc*q5 - dmin*mn ; q5 uniform 0..31 (std 9.23), sc,mn 6-bit.  Keep sc in [24,63], mn ~ 15.5*sc
        d = (sigma / (9.23 * 44.0) * jitter).astype(np.float16)
        out[:, 0:2] = d.view(np.uint8).reshape(nblk, 2)
        sc = rng.integers(24, 64, size=(nblk, 8), dtype=np.uint8)
        mn = rng.integers(20, 44, size=(nblk, 8), dtype=np.uint8)
        # dmin chosen so dmin*mn ~= 15.5*d*sc on average -> zero-mean weights
        dmin = (d.astype(np.float32) * 15.5 * 44.0 / 32.0).astype(np.float16)
        out[:, 2:4] = dmin.view(np.uint8).reshape(nblk, 2)
        out[:, 4:16] = pack_scales_k4(sc, mn)
        out[:, 16:] = rng.integers(0, 256, size=(nblk, 160), dtype=np.uint8)
    elif gtype == GG_Q5_0:
        # w = d*(q5-16), q5 uniform 0..31 (std 9.23)
        out[:, 0:2] = (sigma / 9.23 * jitter).astype(np.float16).view(np.uint8).reshape(nblk, 2)
        out[:, 2:] = rng.integers(0, 256, size=(nblk, 20), dtype=np.uint8)
    elif gtype == GG_Q5_1:
        d = (sigma / 9.23 * jitter).astype(np.float16)
        m = (-15.5 * d.astype(np.float32) * (0.9 + 0.2 * rng.random(nblk, dtype=np.float32))).astype(np.float16)
        out[:, 0:2] = d.view(np.uint8).reshape(nblk, 2)
        out[:, 2:4] = m.view(np.uint8).reshape(nblk, 2)
        out[:, 4:] = rng.integers(0, 256, size=(nblk, 20), dtype=np.uint8)
    elif gtype == GG_Q8_0:
        # w = d*q8, q8 uniform -127..127 (std 73.3)
        out[:, 0:2] = (sigma / 73.3 * jitter).astype(np.float16).view(np.uint8).reshape(nblk, 2)
        out[:, 2:] = rng.integers(-127, 128, size=(nblk, 32), dtype=np.int8).view(np.uint8)
    elif gtype == GG_Q4_K:
        # w = d*sc*q4 - dmin*mn ; q4 uniform 0..15 (std 4.61, mean 7.5)
        d = (sigma / (4.61 * 44.0) * jitter).astype(np.float16)
        out[:, 0:2] = d.view(np.uint8).reshape(nblk, 2)
        sc = rng.integers(24, 64, size=(nblk, 8), dtype=np.uint8)
        mn = rng.integers(20, 44, size=(nblk, 8), dtype=np.uint8)
        dmin = (d.astype(np.float32) * 7.5 * 44.0 / 32.0).astype(np.float16)
        out[:, 2:4] = dmin.view(np.uint8).reshape(nblk, 2)
        out[:, 4:16] = pack_scales_k4(sc, mn)
        out[:, 16:] = rng.integers(0, 256, size=(nblk, 128), dtype=np.uint8)
    elif gtype == GG_Q6_K:
        # w = d*sc*(q6-32); q6 uniform 0..63 (std 18.47), sc int8 in +-[32,127]
        out[:, 0:192] = rng.integers(0, 256, size=(nblk, 192), dtype=np.uint8)
        sc = rng.integers(32, 128, size=(nblk, 16), dtype=np.int16)
        sign = rng.integers(0, 2, size=(nblk, 16), dtype=np.int16) * 2 - 1
        out[:, 192:208] = (sc * sign).astype(np.int8).view(np.uint8)
        d = (sigma / (18.47 * 80.0) * jitter).astype(np.float16)
        out[:, 208:210] = d.view(np.uint8).reshape(nblk, 2)
    else:
        raise ValueError(gtype)
    return out.reshape(rows, -1)


def pack_scales_k4(sc: np.ndarray, mn: np.ndarray) -> np.ndarray:
    """Pack 8 six-bit scales + 8 six-bit mins into 12 bytes (Q4_K/Q5_K layout, inverse of get_scale_min_k4)."""
    n = sc.shape[0]
    q = np.zeros((n, 12), dtype=np.uint8)
    for j in range(4):
        q[:, j] = (sc[:, j] & 63) | ((sc[:, j + 4] >> 4) << 6)
        q[:, j + 4] = (mn[:, j] & 63) | ((mn[:, j + 4] >> 4) << 6)
        q[:, j + 8] = (sc[:, j + 4] & 0xF) | ((mn[:, j + 4] & 0xF) << 4)
    return q


# ----------------------------------------------------------------------------------------------
# numpy dequantisers (independent of gguf-py; tests cross-check both against gguf.quants)
# ----------------------------------------------------------------------------------------------
def dequant(gtype: int, raw: np.ndarray, cols: int) -> np.ndarray:
    rows = raw.shape[0]
    if gtype == GG_F32:
        return raw.view(np.float32).reshape(rows, cols)
    if gtype == GG_F16:
        return raw.view(np.float16).reshape(rows, cols).astype(np.float32)
    per, nb = GG_BLOCK[gtype]
    b = raw.reshape(-1, nb)
    if gtype in (GG_Q4_0, GG_Q4_1):
        d = b[:, 0:2].copy().view(np.float16).astype(np.float32)
        off = 2
        if gtype == GG_Q4_1:
            m = b[:, 2:4].copy().view(np.float16).astype(np.float32)
            off = 4
        qs = b[:, off:]
        q = np.concatenate([qs & 0xF, qs >> 4], axis=1).astype(np.float32)
        w = q * d + m if gtype == GG_Q4_1 else (q - 8.0) * d
        return w.reshape(rows, cols)
    if gtype == GG_Q5_K:
        d = b[:, 0:2].copy().view(np.float16).astype(np.float32)
        dmin = b[:, 2:4].copy().view(np.float16).astype(np.float32)
        s = b[:, 4:16]
        sc = np.empty((b.shape[0], 8), np.float32)
        mn = np.empty((b.shape[0], 8), np.float32)
        for j in range(4):
            sc[:, j] = s[:, j] & 63
            mn[:, j] = s[:, j + 4] & 63
            sc[:, j + 4] = (s[:, j + 8] & 0xF) | ((s[:, j] >> 6) << 4)
            mn[:, j + 4] = (s[:, j + 8] >> 4) | ((s[:, j + 4] >> 6) << 4)
        qh, qs = b[:, 16:48], b[:, 48:176]
        w = np.empty((b.shape[0], 256), np.float32)
        for j in range(4):
            lo = (qs[:, 32 * j:32 * j + 32] & 0xF) + (((qh >> (2 * j)) & 1) << 4)
            hi = (qs[:, 32 * j:32 * j + 32] >> 4) + (((qh >> (2 * j + 1)) & 1) << 4)
            w[:, 64 * j:64 * j + 32] = d * sc[:, 2 * j:2 * j + 1] * lo - dmin * mn[:, 2 * j:2 * j + 1]
            w[:, 64 * j + 32:64 * j + 64] = d * sc[:, 2 * j + 1:2 * j + 2] * hi - dmin * mn[:, 2 * j + 1:2 * j + 2]
        return w.reshape(rows, cols)
    if gtype == GG_Q6_K:
        ql, qh = b[:, 0:128], b[:, 128:192]
        sc = b[:, 192:208].copy().view(np.int8).astype(np.float32)
        d = b[:, 208:210].copy().view(np.float16).astype(np.float32)
        w = np.empty((b.shape[0], 256), np.float32)
        for n in range(2):
            l_, h_ = ql[:, 64 * n:64 * n + 64], qh[:, 32 * n:32 * n + 32]
            q1 = ((l_[:, :32] & 0xF) | (((h_ >> 0) & 3) << 4)).astype(np.int32) - 32
            q2 = ((l_[:, 32:] & 0xF) | (((h_ >> 2) & 3) << 4)).astype(np.int32) - 32
            q3 = ((l_[:, :32] >> 4) | (((h_ >> 4) & 3) << 4)).astype(np.int32) - 32
            q4 = ((l_[:, 32:] >> 4) | (((h_ >> 6) & 3) << 4)).astype(np.int32) - 32
            for k, q in enumerate((q1, q2, q3, q4)):
                s = np.repeat(sc[:, 8 * n + 2 * k:8 * n + 2 * k + 2], 16, axis=1)
                w[:, 128 * n + 32 * k:128 * n + 32 * k + 32] = d * s * q
        return w.reshape(rows, cols)
    raise ValueError(gtype)


# ----------------------------------------------------------------------------------------------
# synthetic vocabulary (SentencePiece-like: 3 specials, 256 byte tokens, scored pieces)
# ----------------------------------------------------------------------------------------------
_SEED_WORDS = (
    "Human Assistant Img ImageContent Give the following image You will be able to see once I provide it "
    "to you Please answer my questions what is this picture color of in a and that with for are on as can "
    "describe photo text there about detail one two three red green blue cat dog llama person table story "
    "tell me how many why where when which who does do not yes no left right top bottom small large"
).split()


def make_vocab(n_vocab: int, seed: int = 7):
    """Return list[(bytes, float score)] of length n_vocab. ids: 0 <unk>, 1 <s>, 2 </s>, 3..258 raw bytes."""
    rng = np.random.default_rng(seed)
    vocab: list[tuple[bytes, float]] = [(b"<unk>", 0.0), (b"<s>", 0.0), (b"</s>", 0.0)]
    for b in range(256):
        vocab.append((bytes([b]), 0.0))
    seen = {v[0] for v in vocab}
    pieces: list[bytes] = []

    def add(p: bytes):
        if p and p not in seen and len(vocab) + len(pieces) < n_vocab:
            seen.add(p)
            pieces.append(p)

    for p in (b"##", b"###", b"<", b">", b"</", b":", b" ", b".", b",", b"?", b"!", b"  "):
        add(p)
    for w in _SEED_WORDS:
        for v in (w, " " + w, w.lower(), " " + w.lower()):
            bs = v.encode()
            # all prefixes so that bigram merging can actually reach the full piece
            for k in range(2, len(bs) + 1):
                add(bs[:k])
    letters = b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789 "
    while len(vocab) + len(pieces) < n_vocab:
        k = int(rng.integers(2, 7))
        add(bytes(rng.choice(list(letters), size=k).tolist()))
    for rank, p in enumerate(pieces):
        # longer pieces merge first within a family; otherwise by rank
        vocab.append((p, -float(rank) * 0.01 - 1.0 + 0.5 * len(p)))
    assert len(vocab) == n_vocab
    return vocab


# ----------------------------------------------------------------------------------------------
# LLaMA ggjt v3
# ----------------------------------------------------------------------------------------------
@dataclass
class LlamaSpec:
    n_vocab: int = 32000
    n_embd: int = 4096
    n_mult: int = 256
    n_head: int = 32
    n_layer: int = 32
    wtype: str = "q4_1"          # type of every 2-D tensor ...
    output_type: str | None = None  # ... except output.weight when given (e.g. "q6_k")
    tok_type: str | None = None
    overrides: dict = field(default_factory=dict)  # name-suffix -> type, e.g. {"attention.wv.weight": "q6_k"}
    seed: int = 2
    vocab: list | None = None  # explicit [(piece bytes, score)] of length n_vocab (e.g. converted from a sentencepiece model); default: make_vocab
    # weight-scale knobs (multipliers on the 1/sqrt(fan_in) default) — see DESIGN.md "synthetic weights / conditioning"
    qk_gain: float = 1.0      # wq, wk
    v_gain: float = 1.0       # wv
    o_gain: float = 0.5       # wo
    ffn_gain: float = 1.0     # w1, w3
    down_gain: float = 0.5    # w2

    @property
    def n_ff(self) -> int:
        return ((2 * (4 * self.n_embd) // 3 + self.n_mult - 1) // self.n_mult) * self.n_mult

    @property
    def n_rot(self) -> int:
        return self.n_embd // self.n_head


LLAMA_7B = dict(n_embd=4096, n_head=32, n_layer=32)
LLAMA_13B = dict(n_embd=5120, n_head=40, n_layer=40)
FTYPE_OF = {"f32": 0, "f16": 1, "q4_0": 2, "q4_1": 3, "q8_0": 7, "q5_0": 8, "q5_1": 9, "q4_k": 15, "q5_k": 17, "q6_k": 18}


def write_llama_ggjt(path: str | Path, spec: LlamaSpec) -> dict:
    """Write a ggjt v3 file; returns {'bytes_per_token': algorithmic weight bytes streamed per decoded token}."""
    rng = np.random.default_rng(spec.seed)
    vocab = spec.vocab if spec.vocab is not None else make_vocab(spec.n_vocab)
    assert len(vocab) == spec.n_vocab
    E, FF = spec.n_embd, spec.n_ff
    sig = 1.0 / np.sqrt(E)
    stats = {"bytes_per_token": 0}

    def ttype(name: str) -> int:
        for suf, t in spec.overrides.items():
            if name.endswith(suf):
                return GG_NAMES[t]
        if name == "output.weight" and spec.output_type:
            return GG_NAMES[spec.output_type]
        if name == "tok_embeddings.weight" and spec.tok_type:
            return GG_NAMES[spec.tok_type]
        return GG_NAMES[spec.wtype]

    with open(path, "wb") as f:
        f.write(struct.pack("<II", 0x67676A74, 3))
        f.write(struct.pack("<7I", spec.n_vocab, E, spec.n_mult, spec.n_head, spec.n_layer, spec.n_rot, FTYPE_OF.get(spec.wtype, 1)))
        for text, score in vocab:
            f.write(struct.pack("<I", len(text)))
            f.write(text)
            f.write(struct.pack("<f", score))

        def tensor(name: str, gtype: int, rows: int, cols: int | None, data: np.ndarray):
            nm = name.encode()
            dims = [rows] if cols is None else [cols, rows]  # ne[0] = cols (contraction dim)
            f.write(struct.pack("<III", len(dims), len(nm), gtype))
            f.write(struct.pack(f"<{len(dims)}I", *dims))
            f.write(nm)
            pad = (-f.tell()) % 32
            f.write(b"\0" * pad)
            data.tofile(f)

        def mat(name: str, rows: int, cols: int, sigma: float, streamed: bool = True):
            gt = ttype(name)
            tensor(name, gt, rows, cols, synth_quant(rng, gt, rows, cols, sigma))
            if streamed:
                stats["bytes_per_token"] += rows * gg_row_bytes(gt, cols)

        def vec(name: str, n: int):
            tensor(name, GG_F32, n, None, (1.0 + 0.05 * rng.standard_normal(n)).astype(np.float32))

        mat("tok_embeddings.weight", spec.n_vocab, E, 1.0, streamed=False)
        vec("norm.weight", E)
        mat("output.weight", spec.n_vocab, E, sig)
        for i in range(spec.n_layer):
            p = f"layers.{i}."
            mat(p + "attention.wq.weight", E, E, sig * spec.qk_gain)
            mat(p + "attention.wk.weight", E, E, sig * spec.qk_gain)
            mat(p + "attention.wv.weight", E, E, sig * spec.v_gain)
            mat(p + "attention.wo.weight", E, E, sig * spec.o_gain)
            vec(p + "attention_norm.weight", E)
            mat(p + "feed_forward.w1.weight", FF, E, sig * spec.ffn_gain)
            mat(p + "feed_forward.w2.weight", E, FF, spec.down_gain / np.sqrt(FF))
            mat(p + "feed_forward.w3.weight", FF, E, sig * spec.ffn_gain)
            vec(p + "ffn_norm.weight", E)
    return stats


# ----------------------------------------------------------------------------------------------
# MiniGPT-4 container
# ----------------------------------------------------------------------------------------------
@dataclass
class VisionSpec:
    n_blocks: int = 39            # EVA ViT-g blocks (whatever the file contains; reference minigpt4.cpp:1886-1899)
    n_qformer_layers: int = 12
    cross_attention_freq: int = 2
    n_embd_llm: int = 4096        # llama_proj out features: 4096 -> 7B, 5120 -> 13B (minigpt4.cpp:1614-1627)
    embed_dim: int = 1408
    mlp_dim: int = 6144
    seed: int = 1
    extra_unused: bool = True     # also write tensors the reference never reads (text-branch FFN), like convert.py does
    fast: bool = False            # bench-size files: synthesise F16 matrices from random bit patterns (10x faster than Gaussian)


def _wstr(f, s: str | bytes):
    b = s.encode() if isinstance(s, str) else s
    f.write(struct.pack("<i", len(b)))
    f.write(b)


def write_minigpt4(path: str | Path, spec: VisionSpec) -> None:
    rng = np.random.default_rng(spec.seed)
    D, FF, HID = spec.embed_dim, spec.mlp_dim, 768

    def w16(rows, cols, sigma=0.02):
        if spec.fast:
            # sign | exponent in {2^-8, 2^-7, 2^-6} (x sigma/0.02 via exponent shift is not needed: |w| ~ 0.004..0.03) | 10 random mantissa bits
            bits = rng.integers(0, 1 << 16, size=(rows, cols), dtype=np.uint16)
            expo = (np.uint16(7) + (bits >> np.uint16(13)) % np.uint16(3)).astype(np.uint16)  # biased exponents 7..9
            out = (bits & np.uint16(0x83FF)) | (expo << np.uint16(10))
            return out.view(np.float16)
        return (rng.standard_normal((rows, cols), dtype=np.float32) * sigma).astype(np.float16)

    def v32(n, mean=0.0, sigma=0.02):
        return (mean + sigma * rng.standard_normal(n)).astype(np.float32)

    models: list[tuple[str, dict[str, np.ndarray]]] = []
    ve: dict[str, np.ndarray] = {}
    ve["cls_token"] = v32(D)
    ve["pos_embed"] = (rng.standard_normal((257, D)) * 0.02).astype(np.float32)
    ve["patch_embed.proj.weight"] = (rng.standard_normal((D, 3, 14, 14)) * 0.02).astype(np.float16)
    ve["patch_embed.proj.bias"] = v32(D)
    for i in range(spec.n_blocks):
        p = f"blocks.{i}."
        ve[p + "norm1.weight"] = v32(D, 1.0)
        ve[p + "norm1.bias"] = v32(D)
        ve[p + "attn.q_bias"] = v32(D)
        ve[p + "attn.v_bias"] = v32(D)
        ve[p + "attn.qkv.weight"] = w16(3 * D, D)
        ve[p + "attn.proj.weight"] = w16(D, D)
        ve[p + "attn.proj.bias"] = v32(D)
        ve[p + "norm2.weight"] = v32(D, 1.0)
        ve[p + "norm2.bias"] = v32(D)
        ve[p + "mlp.fc1.weight"] = w16(FF, D)
        ve[p + "mlp.fc1.bias"] = v32(FF)
        ve[p + "mlp.fc2.weight"] = w16(D, FF)
        ve[p + "mlp.fc2.bias"] = v32(D)
    models.append(("visual_encoder", ve))
    models.append(("ln_vision", {"weight": v32(D, 1.0), "bias": v32(D)}))
    models.append(("query_tokens", {"weight": (rng.standard_normal((32, HID)) * 0.02).astype(np.float32)}))
    qf: dict[str, np.ndarray] = {}
    qf["bert.embeddings.position_ids"] = np.arange(512, dtype=np.float32)
    qf["bert.embeddings.LayerNorm.weight"] = v32(HID, 1.0)
    qf["bert.embeddings.LayerNorm.bias"] = v32(HID)
    for i in range(spec.n_qformer_layers):
        p = f"bert.encoder.layer.{i}."
        for att, kvdim in (("attention", HID), ("crossattention", D)):
            if att == "crossattention" and i % spec.cross_attention_freq != 0:
                continue
            qf[p + att + ".self.query.weight"] = w16(HID, HID)
            qf[p + att + ".self.query.bias"] = v32(HID)
            qf[p + att + ".self.key.weight"] = w16(HID, kvdim)
            qf[p + att + ".self.key.bias"] = v32(HID)
            qf[p + att + ".self.value.weight"] = w16(HID, kvdim)
            qf[p + att + ".self.value.bias"] = v32(HID)
            qf[p + att + ".output.dense.weight"] = w16(HID, HID)
            qf[p + att + ".output.dense.bias"] = v32(HID)
            qf[p + att + ".output.LayerNorm.weight"] = v32(HID, 1.0)
            qf[p + att + ".output.LayerNorm.bias"] = v32(HID)
        if spec.extra_unused and i == 0:
            qf[p + "intermediate.dense.weight"] = w16(16, HID)  # text branch: present in real files, never read
            qf[p + "intermediate.dense.bias"] = v32(16)
        qf[p + "intermediate_query.dense.weight"] = w16(3072, HID)
        qf[p + "intermediate_query.dense.bias"] = v32(3072)
        qf[p + "output_query.dense.weight"] = w16(HID, 3072)
        qf[p + "output_query.dense.bias"] = v32(HID)
        qf[p + "output_query.LayerNorm.weight"] = v32(HID, 1.0)
        qf[p + "output_query.LayerNorm.bias"] = v32(HID)
    models.append(("Qformer", qf))
    models.append(("llama_proj", {"weight": w16(spec.n_embd_llm, HID, 0.05), "bias": v32(spec.n_embd_llm)}))

    config = {"ftype": "f16", "Qformer": {"encoder_width": D, "query_length": 32, "num_hidden_layers": spec.n_qformer_layers,
                                          "hidden_size": HID, "num_attention_heads": 12, "intermediate_size": 3072,
                                          "cross_attention_freq": spec.cross_attention_freq, "layer_norm_eps": 1e-12,
                                          "add_cross_attention": True, "vocab_size": 30523}}
    write_container(path, config, models)


def write_container(path: str | Path, config: dict, models: list[tuple[str, dict[str, np.ndarray]]], file_dtype: int = DT_F16,
                    raw_types: dict[str, tuple[int, list[int]]] | None = None) -> None:
    """Container writer (layout of reference convert.py:56-180 / loader minigpt4.cpp:1478-1596).  float16 / float32 arrays are written
    as F16 / F32 tensors; raw_types["<model>.<tensor>"] = (container dtype, ne[]) marks a uint8 array as ready-made quantised blocks."""
    raw_types = raw_types or {}
    with open(path, "wb") as f:
        f.write(b"ggml")
        f.write(struct.pack("<ii", 1, file_dtype))
        _wstr(f, json.dumps(config))
        for name, tensors in models:
            _wstr(f, name)
            f.write(struct.pack("<i", len(tensors)))
            for tname, arr in tensors.items():
                _wstr(f, tname)
                key = f"{name}.{tname}"
                if key in raw_types:
                    dt, shape = raw_types[key][0], list(raw_types[key][1])
                else:
                    dt, shape = (DT_F16 if arr.dtype == np.float16 else DT_F32), list(arr.shape)[::-1]
                f.write(struct.pack("<i", len(shape)))
                f.write(struct.pack(f"<{len(shape)}i", *shape))
                f.write(struct.pack("<i", dt))
            for tname, arr in tensors.items():
                f.seek((f.tell() + 4095) // 4096 * 4096 if f.tell() % 4096 else f.tell())
                np.ascontiguousarray(arr).tofile(f)


def synth_image(seed: int = 1234) -> np.ndarray:
    """F32 CHW [3,224,224] ~ N(0,1): post-normalisation statistics of a real image."""
    return np.random.default_rng(seed).standard_normal((3, 224, 224)).astype(np.float32)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="write synthetic MiniGPT-4 + LLaMA bins")
    ap.add_argument("--out", default="/tmp/mg4")
    ap.add_argument("--size", default="7b", choices=["7b", "13b", "tiny"])
    ap.add_argument("--wtype", default="q4_1")
    ap.add_argument("--blocks", type=int, default=39)
    a = ap.parse_args()
    out = Path(a.out)
    out.mkdir(parents=True, exist_ok=True)
    dims = dict(n_embd=512, n_head=4, n_layer=2, n_vocab=1024) if a.size == "tiny" else (LLAMA_7B if a.size == "7b" else LLAMA_13B)
    spec = LlamaSpec(wtype=a.wtype, **dims)
    st = write_llama_ggjt(out / f"llama-{a.size}-{a.wtype}.bin", spec)
    write_minigpt4(out / f"minigpt4-{a.size}-f16.bin", VisionSpec(n_blocks=a.blocks, n_embd_llm=spec.n_embd))
    print(st)

"""Host logic (CPU): robustness of the tokenizer and the sampler chain behind minigpt4_begin_chat / minigpt4_end_chat against hostile input -
invalid and truncated UTF-8, very long strings, infinite logits, degenerate sampling parameters.  Child processes, so a crash fails the test."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

TOK = r'''
import sys, tempfile, numpy as np
sys.path.insert(0, sys.argv[1])
import minigpt4_cpp_b200 as m
from minigpt4_cpp_b200 import modelgen as mg
ext = m.B200(m.load_library())
d = tempfile.mkdtemp()
mg.write_llama_ggjt(d + "/l.bin", mg.LlamaSpec(n_vocab=600, n_embd=128, n_head=1, n_layer=1, wtype="f16"))
rng = np.random.default_rng(0)
for it in range(1500):
    L = int(rng.integers(0, 200))
    b = bytes(rng.integers(1, 256, size=L).tolist())
    if rng.random() < 0.3 and L > 0:
        b = b[:-1] + bytes([int(rng.choice([0xC3, 0xE2, 0xF0, 0xFF, 0x80]))])   # multi-byte lead with nothing after it
    ids = ext.host_tokenize(d + "/l.bin", b, bool(rng.integers(0, 2)))
    assert all(0 <= i < 600 for i in ids)
ids = ext.host_tokenize(d + "/l.bin", bytes(rng.integers(1, 256, size=200000).tolist()), True)
assert len(ids) > 1000
print("ok")
'''

SAMP = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1])
import minigpt4_cpp_b200 as m
ext = m.B200(m.load_library())
rng = np.random.default_rng(1)
for it in range(400):
    n = int(rng.choice([1, 2, 5, 100, 1000, 32000]))
    lg = (rng.standard_normal(n) * float(rng.choice([0.01, 1, 30]))).astype(np.float32)
    mode = int(rng.integers(0, 5))
    if mode == 1: lg[:] = lg[0]
    if mode == 2 and n > 2: lg[int(rng.integers(0, n))] = np.inf
    if mode == 3 and n > 2: lg[int(rng.integers(0, n))] = -np.inf
    if mode == 4: lg[:] = -np.inf; lg[int(rng.integers(0, n))] = 0.0
    kw = dict(temp=float(rng.choice([0.0, 1e-6, 0.8, 5.0])), top_k=int(rng.choice([0, 1, 40, 10 ** 6, -5])), top_p=float(rng.choice([0.0, 0.01, 0.9, 1.0, 2.0])),
              tfs_z=float(rng.choice([1.0, 0.5, 0.0])), typical_p=float(rng.choice([1.0, 0.5, 0.0])), mirostat=int(rng.choice([0, 0, 1, 2])))
    out = ext.host_sample(lg, seed=int(rng.integers(0, 1000)), n_draws=4, **kw)
    assert ((out >= 0) & (out < n)).all(), (n, kw, out)
print("ok")
'''


def _child(tmp_path, code):
    f = tmp_path / "child.py"
    f.write_text(code)
    r = subprocess.run([sys.executable, str(f), str(ROOT)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), (r.returncode, r.stdout[-200:], r.stderr[-400:])


def test_tokenizer_survives_hostile_bytes(tmp_path):
    _child(tmp_path, TOK)


def test_sampler_survives_degenerate_logits_and_parameters(tmp_path):
    _child(tmp_path, SAMP)

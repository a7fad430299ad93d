"""Host logic (CPU): the file readers behind minigpt4_model_load (ggjt v3 and the MiniGPT-4 container; reference loaders
minigpt4.cpp:1478-1596 and llama.cpp's) must reject damaged files with an error code — never read out of bounds.  Random
truncations, byte flips in the header region and wild 32-bit fields; each batch runs in a child process so that a crash is a test failure
rather than the end of the test run."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

CHILD = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[4])
import minigpt4_cpp_b200 as m
ext = m.B200(m.load_library())
kind, path, seed = sys.argv[1], sys.argv[2], int(sys.argv[3])
data = bytearray(open(path, "rb").read())
rng = np.random.default_rng(seed)
codes = set()
for it in range(60):
    b = bytearray(data)
    mode = int(rng.integers(0, 3))
    if mode == 0:
        b = b[:int(rng.integers(0, len(b)))]
    elif mode == 1:
        for _ in range(int(rng.integers(1, 8))):
            b[int(rng.integers(0, min(len(b), 4096)))] = int(rng.integers(0, 256))
    else:
        i = int(rng.integers(0, min(len(b), 2048)))
        b[i:i + 4] = int(rng.integers(0, 2 ** 31)).to_bytes(4, "little")
    p = path + ".fz"
    open(p, "wb").write(b)
    if kind == "ggjt":
        codes.add(ext.L.minigpt4_b200_host_inspect_ggjt(p.encode(), None, None, None, None))
    else:
        codes.add(ext.L.minigpt4_b200_host_inspect_container(p.encode(), None, None, None))
print("codes", sorted(codes))
'''


@pytest.mark.parametrize("kind", ["ggjt", "container"])
def test_damaged_files_are_rejected_not_crashed_on(mg, tmp_path, kind):
    if kind == "ggjt":
        path = str(tmp_path / "l.bin")
        mg.write_llama_ggjt(path, mg.LlamaSpec(n_vocab=300, n_embd=128, n_head=1, n_layer=1, wtype="q4_1"))
        allowed = {0, 4}          # None, LoadLanguageModel
    else:
        path = str(tmp_path / "v.bin")
        mg.write_minigpt4(path, mg.VisionSpec(n_blocks=1, n_qformer_layers=1, n_embd_llm=4096))
        allowed = {0, 1, 2, 3}    # None, LoadModelFileHeader, LoadModelFileVersion, LoadModelMiniGPT4DataType
    child = tmp_path / "child.py"
    child.write_text(CHILD)
    for seed in (0, 1):
        r = subprocess.run([sys.executable, str(child), kind, path, str(seed), str(ROOT)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (r.returncode, r.stderr[-400:])
        codes = set(eval(r.stdout.strip().split("codes", 1)[1]))
        assert codes <= allowed and len(codes) >= 2, codes

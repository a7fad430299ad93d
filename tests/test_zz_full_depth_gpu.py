"""The benchmarked vision configuration itself (EVA ViT-g/14 with 39 blocks + 12 Q-Former layers + llama_proj, reference
minigpt4.cpp:2094-2363) against the CPU oracle through minigpt4_encode_image - the same comparison bench.py prints as
`parity.vision_rel_err_full_depth`, as a test.  The 2 GB synthetic container is shared with bench.py (same path), so one of the two generates it.
(File name sorts last: the heavy case runs after the rest of the suite.)"""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def test_full_depth_vision_graph_matches_oracle(lib, ext, orc, mg, tmp_path):
    import bench
    vis = bench.model_dir() / "minigpt4-7b-f16-b39.bin"
    if not vis.exists():   # (written under another name and renamed: an interrupted run must not leave a truncated file where bench.py looks)
        import os
        part = vis.with_name(vis.name + f".part{os.getpid()}")
        mg.write_minigpt4(part, mg.VisionSpec(n_blocks=39, n_embd_llm=4096, fast=True))
        os.replace(part, vis)
    llm = tmp_path / "llama-4096.bin"
    mg.write_llama_ggjt(llm, mg.LlamaSpec(n_vocab=2048, n_embd=4096, n_head=32, n_layer=2, wtype="q4_1"))
    c = lib.minigpt4_model_load(str(vis), str(llm), 1, 1, 256, 8, 0)
    assert c.ptr
    img = mg.synth_image()
    got = ext.encode_array(c, img)
    want = orc.OracleEngine(str(vis), None).encode_image(img)
    assert got.shape == want.shape == (32, 4096)
    assert np.isfinite(got).all()
    assert rel_err(got, want) < 1e-2, rel_err(got, want)   # (measured 3.8e-4)
    lib.minigpt4_free(c)

"""World-size-2 (gloo, CPU) check of the tensor-parallel plan used by LlamaDevice::load:
row-split matmuls (wo, down) are column-sharded on quant-block boundaries, each rank computes a partial with the
oracle's integer-dot mul_mat on ITS shard of weights and activations, partials are summed with an all-reduce, and the
result equals the unsharded product up to float summation order — because activation quantisation is per 32/256-block
and happens on the full (post-reduction) vector.  Column-split matmuls (qkv, gate/up) are sharded by output rows."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank: int, world: int, port: int, q):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    from minigpt4_cpp_b200 import modelgen as mg
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = {}
        for name, gt, blk in (("q4_1", 3, 32), ("q5_k", 13, 256), ("f16", 1, 1)):
            rows, cols = 64, 1024
            rng = np.random.default_rng(123)  # same weights on every rank
            raw = mg.synth_quant(rng, gt, rows, cols, 0.02)
            x = rng.standard_normal((3, cols)).astype(np.float32)
            full = orc.mul_mat(gt, raw, rows, cols, x)
            # row-parallel ("wo"/"down"): column shard
            cl = cols // world
            assert cl % blk == 0
            rb = mg.gg_row_bytes(gt, cols) // world
            part = orc.mul_mat(gt, np.ascontiguousarray(raw[:, rank * rb:(rank + 1) * rb]), rows, cl, np.ascontiguousarray(x[:, rank * cl:(rank + 1) * cl]))
            t = torch.from_numpy(part.copy())
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            out[name + "_rowpar"] = float(np.abs(t.numpy() - full).max() / np.abs(full).max())
            # column-parallel ("qkv"/"gate-up"): output-row shard, all-gather reproduces the full result bit for bit
            rl = rows // world
            mine = orc.mul_mat(gt, np.ascontiguousarray(raw[rank * rl:(rank + 1) * rl]), rl, cols, x)
            parts = [torch.zeros(3, rl) for _ in range(world)]
            dist.all_gather(parts, torch.from_numpy(mine.copy()))
            out[name + "_colpar_exact"] = bool(np.array_equal(torch.cat(parts, dim=1).numpy(), full))
        # the NCCL-id style bootstrap used by bench.py: rank 0 creates 128 bytes, everyone must end up with them
        uid = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
        dist.broadcast(uid, 0)
        out["uid_ok"] = bool((uid == torch.arange(128, dtype=torch.uint8)).all())
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_tensor_parallel_plan_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank in (0, 1):
        o = res[rank]
        assert o["uid_ok"]
        for name in ("q4_1", "q5_k", "f16"):
            assert o[name + "_rowpar"] < 1e-6, (name, o)
            assert o[name + "_colpar_exact"], name


def test_split_alignment_rules():
    """7B/13B shard sizes vs quant-block sizes (why K-quant TP splits can be refused at load)."""
    for n_embd, n_head, n_ff in ((4096, 32, 11008), (5120, 40, 13824)):
        for world in (2, 4, 8):
            el, ffl = n_embd // world, n_ff // world
            assert el % 32 == 0 and (n_ff % world or ffl % 32 == 0)
    assert (5120 // 8) % 256 != 0  # 13B Q5_K at TP=8: wo column shard is not super-block aligned -> load refuses it
    assert (11008 // 2) % 256 != 0  # 7B K-quant FFN split at TP=2 is not aligned either

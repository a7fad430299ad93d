"""Tensor-parallel check (run under torchrun, one rank per GPU): TP=N logits vs the CPU oracle on a tiny model and on 7B.
TP differs from 1 GPU only by the float order of the partial sums (all-reduce after wo / down), so the comparison is by
tolerance (north_star bar 1e-2) plus greedy-token agreement."""
import ctypes, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import minigpt4_cpp_b200 as m
from minigpt4_cpp_b200 import modelgen as mg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = m.load_library(); ext = m.B200(lib)
ext.L.minigpt4_b200_set_device(local)
uid = np.zeros(128, np.uint8)
if rank == 0:
    ext.L.minigpt4_b200_tp_unique_id(uid.ctypes.data_as(ctypes.c_void_p))
t = torch.from_numpy(uid).cuda(); dist.broadcast(t, 0); uid = t.cpu().numpy()
ext.L.minigpt4_b200_tp_configure(rank, world, uid.ctypes.data_as(ctypes.c_void_p))
d = "/dev/shm/tpcheck"; os.makedirs(d, exist_ok=True)
out = {}
for name, spec in (("tiny_q4_1", mg.LlamaSpec(n_vocab=1024, n_embd=1024, n_head=8, n_layer=4, wtype="q4_1")),
                   ("tiny_f16", mg.LlamaSpec(n_vocab=1024, n_embd=1024, n_head=8, n_layer=4, wtype="f16"))):
    p = f"{d}/{name}.bin"
    if rank == 0:
        mg.write_llama_ggjt(p, spec)
    dist.barrier()
    c = ext.llm_load(p, n_ctx=128)
    ids = list(range(5, 26))
    ext.eval_tokens(c, ids)
    lg = ext.logits(c)
    g = []
    for _ in range(8):
        tid = ext.greedy_id(c); g.append(tid); ext.eval_tokens(c, [tid])
    if rank == 0:
        from oracle import oracle as orc
        e = orc.OracleEngine(None, p, n_ctx=128)
        e.eval_tokens(ids)
        lc = e.logits.copy()
        cg = [e.end_chat_greedy()[0] for _ in range(8)]
        out[name] = {"logits_rel_err": float(np.abs(lg - lc).max() / np.abs(lc).max()), "greedy_match": g == cg, "tp": ext.stats(c).tp_world}
    allg = [None] * world
    dist.all_gather_object(allg, g)
    assert all(a == allg[0] for a in allg), "ranks disagree on greedy ids"
    lib.minigpt4_free(c)
if rank == 0:
    print(json.dumps({"tp_check": out, "world": world}))
dist.destroy_process_group()

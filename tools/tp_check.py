"""Tensor-parallel check (run under torchrun, one rank per GPU): TP=N logits vs the CPU oracle on tiny models (per-op path: F16, K-quants ...)
and on a 7B-wide Q4_1 model (the megakernel with the in-kernel peer all-reduce).  TP differs from one GPU only by the float order of the
partial sums (all-reduce after wo / down), so the comparison is by tolerance (north_star bar 1e-2) plus greedy-token agreement; all ranks must
agree with each other bit for bit (every rank sums the partials in rank order)."""
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


def tp_configure():   # a NCCL unique id serves ONE communicator: a fresh id for every engine that is loaded
    uid = np.zeros(128, np.uint8)
    if rank == 0:
        ext.L.minigpt4_b200_tp_unique_id(uid.ctypes.data_as(ctypes.c_void_p))
    t = torch.from_numpy(uid).cuda(); dist.broadcast(t, 0)
    ext.L.minigpt4_b200_tp_configure(rank, world, t.cpu().numpy().ctypes.data_as(ctypes.c_void_p))


d = "/dev/shm/tpcheck"; os.makedirs(d, exist_ok=True)
out = {}
cases = (("tiny_q4_1", mg.LlamaSpec(n_vocab=1024, n_embd=1024, n_head=8, n_layer=4, wtype="q4_1"), 21, 8),
         ("tiny_f16", mg.LlamaSpec(n_vocab=1024, n_embd=1024, n_head=8, n_layer=4, wtype="f16"), 21, 8),
         ("wide_q4_1_megakernel", mg.LlamaSpec(n_vocab=2048, n_embd=4096, n_head=32, n_layer=2, wtype="q4_1"), 40, 32),
         ("wide_q5_k_megakernel", mg.LlamaSpec(n_vocab=2048, n_embd=4096, n_head=32, n_layer=2, n_mult=512, wtype="q5_k"), 40, 32))   # n_ff 11264: shards of whole super-blocks for 2 and 4 ranks
only = os.environ.get("TP_CHECK_ONLY")   # run one case by name
for name, spec, n_prompt, n_gen in cases:
    if only and name != only:
        continue
    p = f"{d}/{name}.bin"
    if rank == 0:
        mg.write_llama_ggjt(p, spec)
    dist.barrier()
    tp_configure()
    c = ext.llm_load(p, n_ctx=256)
    ids = list(range(5, 5 + n_prompt))
    ext.eval_tokens(c, ids)
    lg = ext.logits(c)
    g = []
    for _ in range(n_gen):
        tid = ext.greedy_id(c); g.append(tid); ext.eval_tokens(c, [tid])
    ch, _ = ext.decode_chain(c, 8)   # chained greedy steps: every rank launches its own graph, the kernels meet in the all-reduce
    st = ext.stats(c)
    ar_us, peer = ext.tp_time_allreduce(c, 32)
    if rank == 0:
        from oracle import oracle as orc
        e = orc.OracleEngine(None, p, n_ctx=256)
        e.eval_tokens(ids)
        lc = e.logits.copy()
        cg = [e.end_chat_greedy()[0] for _ in range(n_gen)]
        n_same = next((i for i, (a, b) in enumerate(zip(g, cg)) if a != b), len(g))
        out[name] = {"logits_rel_err_vs_oracle": float(np.abs(lg - lc).max() / np.abs(lc).max()), "leading_greedy_ids_equal_to_oracle": n_same, "of": n_gen,
                     "tp": st.tp_world, "decode_megakernel": st.decode_megakernel, "peer_allreduce": peer, "allreduce_us": ar_us}
    allg = [None] * world
    dist.all_gather_object(allg, (g, ch.tolist(), lg.tobytes()))
    assert all(a == allg[0] for a in allg), "ranks disagree (ids / chained ids / logits bits)"
    lib.minigpt4_free(c)
if rank == 0:
    print(json.dumps({"tp_check": out, "world": world}))
    assert all(v["logits_rel_err_vs_oracle"] < 1e-2 for v in out.values()), out
dist.destroy_process_group()

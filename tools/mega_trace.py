"""Per-op timeline of the decode megakernel (needs MINIGPT4_B200_MEGA_TRACE=1): prints where a token's time goes."""
import os, sys, json
NOTRACE = bool(os.environ.get("NOTRACE"))
if not NOTRACE: os.environ["MINIGPT4_B200_MEGA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import minigpt4_cpp_b200 as m
import bench
lib = m.load_library(); ext = m.B200(lib)
from minigpt4_cpp_b200 import modelgen as mg
NL = int(os.environ.get("TRACE_LAYERS", "4"))  # per-layer numbers do not depend on depth; 4 layers keep generation fast
NV = int(os.environ.get("TRACE_VOCAB", "32000"))  # a huge vocabulary turns the token into one long barrier-free weight stream (the output matvec)
llm = str(bench.model_dir() / f"llama-7bwide-{NL}l-v{NV}-q4_1.bin")
if not os.path.exists(llm):
    mg.write_llama_ggjt(llm, mg.LlamaSpec(wtype="q4_1", n_vocab=NV, n_embd=4096, n_head=32, n_layer=NL))
ctx = ext.llm_load(llm, n_ctx=2048)
rows = np.random.default_rng(0).standard_normal((32, 4096)).astype(np.float32)
ext.eval_embd(ctx, rows)
ids, ms = ext.decode_chain(ctx, 128)
print("chain ms/token", ms / 128, "(trace off)" if NOTRACE else "(trace on)", f"| {ext.stats(ctx).llm_weight_bytes_per_token / (ms / 128) * 1e-6:.0f} GB/s of weights")
if NOTRACE: sys.exit(0)
tr = ext.mega_trace(ctx).astype(np.float64)  # last launch
names = {0: "embed", 1: "qkv", 2: "attn", 3: "wo", 4: "gate_up", 5: "down", 6: "output", 7: "final"}
kinds = [0] + [1, 2, 3, 4, 5] * NL + [6, 7]
mhz = 1965.0
sub_agg = {}
for c in range(2):
    t = tr[c]
    tot = (t[-1, 1] - t[0, 0]) / mhz
    print(f"CTA {'0' if c == 0 else 'G-1'}: total {tot:.1f} us")
    agg = {}
    for i, k in enumerate(kinds):
        start, bar, staged, done, wwait, wdot, wunits, wepi = t[i][:8]
        nxt = t[i + 1, 0] if i + 1 < len(kinds) else t[i, 1]
        a = agg.setdefault(names[k], [0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += (bar - start) / mhz
        if staged: a[2] += (staged - bar) / mhz; a[3] += (done - staged) / mhz
        a[4] += (nxt - start) / mhz; a[5] += wwait / mhz; a[6] += wdot / mhz; a[7] += wunits; a[8] += wepi / mhz
        if t.shape[1] >= 12:   # generation 6 sub-stamps
            sub = sub_agg.setdefault((c, names[k]), [0.0] * 5)
            if names[k] == "attn" and t[i][8]: sub[0] += (t[i][8] - bar) / mhz; sub[1] += (t[i][9] - t[i][8]) / mhz; sub[2] += (t[i][10] - t[i][9]) / mhz; sub[3] += (t[i][11] - t[i][10]) / mhz; sub[4] += 1
            elif t[i][9] and staged:
                if t[i][8]: sub[0] += (t[i][8] - bar) / mhz; sub[1] += (t[i][9] - t[i][8]) / mhz
                else: sub[1] += (t[i][9] - bar) / mhz
                sub[2] += (staged - t[i][9]) / mhz; sub[4] += 1
    for k, a in agg.items():
        print(f"  {k:8s} n={a[0]:3d} barrier {a[1]/a[0]:6.2f} us  stage {a[2]/a[0]:6.2f} us  consume {a[3]/a[0]:6.2f} us  total/op {a[4]/a[0]:6.2f} us  sum {a[4]:8.1f} us"
              f" | warp0: fill-wait {a[5]/a[0]:5.2f} us, dot {a[6]/a[0]:5.2f} us, epilogue {a[8]/a[0]:5.2f} us, {a[7]/a[0]:4.1f} units/op")
for (c, name), sub in sub_agg.items():
    if not sub[4]: continue
    n = sub[4]
    if name == "attn": print(f"  CTA {'0' if c == 0 else 'G-1'} attn: scores+max {sub[0]/n:.2f} us | exp+sum {sub[1]/n:.2f} | probabilities {sub[2]/n:.2f} | P.V+tree {sub[3]/n:.2f}")
    else: print(f"  CTA {'0' if c == 0 else 'G-1'} {name} staging: load + sum of squares {sub[0]/n:.2f} us | quantise {sub[1]/n:.2f} | final CTA sync {sub[2]/n:.2f}")

"""Model check (CPU, pure Python) of the flag-in-data exchange protocol used by csrc/llama_mega_ll.cuh.

The CUDA variant removes every grid barrier inside a layer and relies on two claims (its header):
  (1) no deadlock: every element a consumer spins on is eventually written with the tag it expects;
  (2) no write-after-read hazard: a cell is never overwritten with version n+1 while some consumer still needs version n
      (if it were, that consumer would spin forever on a tag that is gone - so (2) failing shows up as (1) failing).
This script executes the same op program - EMBED, then per layer QKV -> ATTN -> WO -> GATE/UP -> DOWN, then OUTPUT - on G simulated CTAs with
the kernel's work partition (contiguous row-pair units per CTA, one head per CTA for attention, every staging pass gathers the WHOLE input
vector element by element), under a random scheduler that advances ONE element access of ONE CTA at a time (the most adversarial interleaving
a real GPU could produce), for several launches in a row, and checks that every gather returns exactly the version it was meant to read.

    python tools/ll_protocol_sim.py [seeds]
"""
from __future__ import annotations

import random
import sys


def unit_begin(cta: int, n: int, G: int) -> int:
    return cta * n // G


def program(G: int, H: int, L: int, E: int, FF: int, bug: str = ""):
    """Per-CTA generator of micro-steps.  A step is ('w', vec, idx, tag, payload) or ('r', vec, idx, tag, expected payload)."""
    hd = E // H  # elements per head

    def cta_prog(cta: int, launches: int):
        for seq in range(launches):
            tag0 = (seq << 10) + 1
            oi = 0
            # EMBED: x spread over the grid
            t_embed = tag0 + oi
            for i in range(cta, E, G):
                yield ("w", "x", i, t_embed, ("x", seq, 0))
            tag_x, ver_x = t_embed, 0
            oi += 1
            for layer in range(L):
                # ---- QKV: gather x, write own rows of q | k | v (3E rows = 3E/2 units)
                for i in range(E):
                    yield ("r", "x", i, tag_x, ("x", seq, ver_x))
                t_qkv = tag0 + oi
                lo, hi = unit_begin(cta, 3 * E // 2, G), unit_begin(cta + 1, 3 * E // 2, G)
                for su in range(lo, hi):
                    r0 = 2 * su
                    part, rr = r0 // E, r0 % E
                    if part == 0:
                        yield ("w", "q", rr, t_qkv, ("q", seq, layer)); yield ("w", "q", rr + 1, t_qkv, ("q", seq, layer))
                    elif part == 1:
                        yield ("w", "kcur", rr // 2, t_qkv, ("k", seq, layer))
                    else:
                        yield ("w", "vcur", rr // 2, t_qkv, ("v", seq, layer))
                oi += 1
                # ---- ATTN: head CTAs gather their q / k / v slices, write att
                t_att = tag0 + oi
                if cta < H:
                    h = cta
                    for i in range(h * hd, (h + 1) * hd):
                        yield ("r", "q", i, t_qkv, ("q", seq, layer))
                    for i in range(h * hd // 2, (h + 1) * hd // 2):
                        yield ("r", "kcur", i, t_qkv, ("k", seq, layer))
                        yield ("r", "vcur", i, t_qkv, ("v", seq, layer))
                    for i in range(h * hd, (h + 1) * hd):
                        yield ("w", "att", i, t_att, ("att", seq, layer))
                oi += 1
                # ---- WO: gather att, x' = x + wo(att) on own rows (reads its own old x rows without waiting: they are final)
                for i in range(hd if bug == "wo_partial_gather" else E):   # (negative control: a WO that only waits for head 0 breaks claim (2))
                    yield ("r", "att", i, t_att, ("att", seq, layer))
                t_wo = tag0 + oi
                lo, hi = unit_begin(cta, E // 2, G), unit_begin(cta + 1, E // 2, G)
                for su in range(lo, hi):
                    for r in (2 * su, 2 * su + 1):
                        yield ("r", "x", r, tag_x, ("x", seq, ver_x))          # residual (plain read in the kernel; here also version-checked)
                        yield ("w", "x", r, t_wo, ("x", seq, ver_x + 1))
                tag_x, ver_x = t_wo, ver_x + 1
                oi += 1
                # ---- GATE/UP: gather x, write own act elements
                for i in range(E):
                    yield ("r", "x", i, tag_x, ("x", seq, ver_x))
                t_gu = tag0 + oi
                lo, hi = unit_begin(cta, FF, G), unit_begin(cta + 1, FF, G)
                for su in range(lo, hi):
                    yield ("w", "act", su, t_gu, ("act", seq, layer))
                oi += 1
                # ---- DOWN: gather act, x'' = x' + down(act) on own rows
                for i in range(FF):
                    yield ("r", "act", i, t_gu, ("act", seq, layer))
                t_dn = tag0 + oi
                lo, hi = unit_begin(cta, E // 2, G), unit_begin(cta + 1, E // 2, G)
                for su in range(lo, hi):
                    for r in (2 * su, 2 * su + 1):
                        yield ("r", "x", r, tag_x, ("x", seq, ver_x))
                        yield ("w", "x", r, t_dn, ("x", seq, ver_x + 1))
                tag_x, ver_x = t_dn, ver_x + 1
                oi += 1
            # ---- OUTPUT: gather x; then the one grid barrier + FINAL
            for i in range(E):
                yield ("r", "x", i, tag_x, ("x", seq, ver_x))
            yield ("barrier", seq)

    return cta_prog


def run(seed: int, G: int = 7, H: int = 3, L: int = 2, E: int = 24, FF: int = 40, launches: int = 3, bug: str = "", skew: bool = False) -> int:
    rng = random.Random(seed)
    weights = [1.0] * G
    mem = {v: [(0, None)] * n for v, n in (("x", E), ("q", E), ("att", E), ("act", FF), ("kcur", E // 2), ("vcur", E // 2))}
    prog = program(G, H, L, E, FF, bug)
    gens = [prog(c, launches) for c in range(G)]
    pending = [next(g) for g in gens]          # the step each CTA is trying to perform
    at_barrier: dict[int, set] = {}
    done = [False] * G
    steps = idle = 0
    while not all(done):
        if skew and (steps + idle) % 997 == 0:   # re-draw who is fast and who is nearly stalled: long stretches of extreme imbalance
            weights = [rng.choice((0.001, 0.05, 1.0, 20.0)) for _ in range(G)]
        c = rng.choices(range(G), weights)[0] if skew else rng.randrange(G)
        if done[c]:
            idle += 1
            if idle > 20000 * G and skew:
                weights = [1.0] * G
            continue
        st = pending[c]
        progressed = False
        if st[0] == "w":
            _, v, i, tag, payload = st
            mem[v][i] = (tag, payload)
            progressed = True
        elif st[0] == "r":
            _, v, i, tag, want = st
            have_tag, have = mem[v][i]
            if have_tag == tag:
                assert have == want, f"seed {seed}: CTA {c} read {v}[{i}] = {have}, wanted {want}"
                progressed = True
        else:  # grid barrier in front of FINAL: wait until every CTA has arrived for this launch
            arrived = at_barrier.setdefault(st[1], set())
            arrived.add(c)
            progressed = len(arrived) == G
        if progressed:
            steps += 1; idle = 0
            try:
                pending[c] = next(gens[c])
            except StopIteration:
                done[c] = True
        else:
            idle += 1
            if idle > (200000 if skew else 200) * G:  # every CTA has been offered a turn many times without any progress
                blocked = {k: pending[k] for k in range(G) if not done[k]}
                raise AssertionError(f"seed {seed}: deadlock after {steps} steps; pending = {blocked}")
    return steps


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    total = 0
    for s in range(n):
        total += run(s)
        total += run(3000 + s, skew=True)
        total += run(4000 + s, G=9, H=2, L=2, E=16, FF=24, launches=2, skew=True)
        total += run(1000 + s, G=5, H=5, L=1, E=20, FF=12, launches=2)     # every CTA is a head CTA
        total += run(2000 + s, G=9, H=2, L=3, E=16, FF=48, launches=2)     # few heads, many idle CTAs during attention
    caught = 0
    for s in range(n):   # negative control: the checker must catch a protocol that violates the full-gather rule
        try:
            run(5000 + s, bug="wo_partial_gather", skew=True)
        except AssertionError:
            caught += 1
    assert caught > 0, "negative control was never caught: the model check has no teeth"
    print(f"negative control (WO waits for one head only) caught in {caught} of {n} skewed schedules")
    print(f"ok: {5 * n} random schedules, {total} element accesses, no deadlock, every gather saw the version it was meant to see")

// llama_mega5.cuh — generation 5 of the persistent decode megakernel (one launch per generated token).
//
// What the round-1 kernel (llama_mega.cuh, "v4") measured: 45 us per layer against 19.2 us of HBM time; the weight stream itself is fine
// (DRAM traffic = algorithmic bytes) but (1) the consumer warps drain a FULL ring at only ~1.06x the HBM rate on n_embd-wide matrices
// (8 shared-memory loads per 2 quant blocks: the staged activations are re-read for every row pair), so time lost in a stall is never
// caught up; (2) every op boundary is a grid barrier + reload (5 per layer, ~2.5 us + ~2.3 us each); (3) the attention op runs on n_head
// CTAs between two grid barriers while 116 CTAs idle.  This generation changes the three of them:
//   * REGISTER-RESIDENT ACTIVATIONS for n_embd-wide inputs (qkv, wo, gate/up, output): lane l of every consumer warp keeps the Q8 blocks
//     l, l+32, ... of the staged vector in registers (40 registers at n_embd 4096), so a row pair costs 8 + 8 shared-memory loads of
//     WEIGHTS only and the canonical per-lane block order (oracle.cpp) is kept exactly.
//   * FLAG-IN-DATA EXCHANGE: the vectors ops exchange (x, q, current K/V, att, act) are {payload, tag} pairs written and read as ONE 64-bit
//     access, tag = (launch sequence << 10) + op index + 1.  A consumer's load IS its barrier; there is no grid barrier inside a layer
//     (one remains in front of the arg-max).  The hazard argument is in the comment of `decode_megakernel5`.
//   * ATTENTION OUTPUT IS EXCHANGED QUANTISED: the head's CTA quantises its 128 outputs to the Q8 blocks `wo` consumes (the blocks are
//     local to a head) and publishes 40 tagged words per head; the `wo` staging of every CTA is then a 10 KB gather instead of a 32 KB
//     gather + quantisation.
//   * GROUP SLOTS: a ring slot holds SEVERAL consecutive units (7B: 4 row pairs of an n_embd-wide matrix = 20 480 B, or 3 rows of the n_ff-wide
//     one = 20 640 B, in a 20 736 B slot) filled by ONE cp.async.bulk.  Measured on a barrier-free stream (1 layer + 200 000-row output matrix):
//     with one 5 KB copy per row pair the single producer thread tops out near 4.4 TB/s (its per-copy cost: try_wait + expect_tx + issue) and a
//     quarter of every 6.9 KB slot stayed empty; group slots cut the producer's work per byte by 4 and use 99 % of the ring.
// Arithmetic and every float reduction order are those of llama_kernels.cuh / oracle.cpp: logits stay bit-identical.
#pragma once
#include "llama_mega.cuh"

namespace mg4 {
namespace mk5 {
using namespace k;
using mk::OP_EMBED; using mk::OP_QKV; using mk::OP_ATTN; using mk::OP_WO; using mk::OP_GATEUP; using mk::OP_DOWN; using mk::OP_OUTPUT; using mk::OP_FINAL;
using mk::kConsumerWarps; using mk::kConsumerThreads; using mk::kMegaThreads;
using mk::consumer_sync; using mk::smem_addr; using mk::mb_init; using mk::mb_expect_tx; using mk::mb_arrive; using mk::mb_wait; using mk::bulk_g2s;
using mk::prefetch_l2; using mk::prefetch_kv_head; using mk::unit_begin; using mk::grid_barrier; using mk::q4_block_idot;
using mk::kNormItems; using mk::kPlainItems;

struct LLf { float v; unsigned tag; };  // 8 bytes, 8-byte aligned: one 64-bit access

struct Op5 {                 // one op of the token program, 32 bytes (the program lives in shared memory: 5 n_layer + 3 entries)
    int cols, n_su;          // input width; units of the op over the whole grid (a unit = a row pair; `down`: ONE n_ff-wide row)
    unsigned short row_bytes, layer;
    unsigned char kind, rpu, n_warps, upg;   // rpu = rows per unit (2 | 1); upg = units per ring slot: ONE bulk copy brings upg consecutive units
    const unsigned char *w;  // row-packed Q4 weights (null for non-matvec ops)
    const float *norm_w;
};
static_assert(sizeof(Op5) == 32, "Op5 is copied as two uint4");

struct Params {
    const Op5 *ops; int n_ops;
    int n_slots, slot_bytes, ff_bytes, e_bytes;  // shared memory: [ring][ff: staged `down` / `wo` input, attention scratch][e0: staged qkv / gate-up / output input][mbarriers][ops]
    int E, FF, n_head, n_ctx, n_vocab;
    float kq_scale;
    LLf *x, *q, *act;            // [E], [E], [FF] tagged floats
    LLf *kcur, *vcur;            // [E/2] each: the current position's K / V as tagged half2 (the F16 cache rows have no room for a tag)
    LLf *att;                    // quantised attention output: [E/32 * 8] words of 4 int8 | [E/32] d | [E/32] s, all tagged
    unsigned *seq;               // launches so far (advanced by OP_FINAL); part of every tag
    unsigned *done;              // [n_ops] monotonic completion counters: CTAs that finished op i (all launches).  Only a HINT that tells pollers
                                 // when to look (no fence orders it against the data); readiness itself is the tag in every element
    volatile unsigned *dbg;      // pinned host word: reason code of a spin-guard trap (readable after the launch failed)
    float *logits;
    __half *kcache, *vcache;
    const float2 *rope; const __half *tab_exp, *tab_silu;
    const unsigned char *tok; int tok_type; size_t tok_row_bytes;
    DeviceState *state; unsigned *barrier;
    int upg_max;                 // arrival count of the slots' "empty" barriers = the largest upg of the program
    int flags;                   // bit 0: request the head's K/V history into L2 while the qkv weights are consumed
                                 // bit 2: CTA barrier BEFORE staging too (polling starts when the whole CTA is done with the previous op)
    long long *trace;            // optional [2 CTAs][n_ops][8] clock64 stamps (layout of tools/mega_trace.py)
};

// ---- staged activation layout in shared memory (Q8_0 / Q8_1): [lo plane cols/2][64 B pad][hi plane cols/2][d: nb floats][s: nb floats].
// The pad puts the two 16-byte planes of a block 16 banks apart, so the staging stores of a warp (lanes 0-3 -> lo, 4-7 -> hi) do not collide.
__host__ __device__ inline int act5_hi(int cols) { return cols / 2 + 64; }
__host__ __device__ inline int act5_d(int cols) { return cols + 64; }
__host__ __device__ inline size_t act5_bytes(int cols) { return (size_t)cols + 64 + (size_t)cols / 32 * 8; }

__device__ __forceinline__ void ll_store(LLf *p, float v, unsigned tag) {
    const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ void ll_store_bits(LLf *p, unsigned bits, unsigned tag) {
    const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)bits;
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const LLf *p) {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
// Every spin is bounded: a protocol bug must end in a trapped launch (an error the host reports), never in a hung GPU.  The reason code goes to
// a pinned host word first (device memory is unreadable after a trap).
constexpr unsigned kSpinLimit = 1u << 22;
__device__ unsigned *g_dbg_word = nullptr;
__device__ __forceinline__ void spin_guard(unsigned &spins, unsigned code) {
    if (++spins > kSpinLimit) { if (g_dbg_word) { *(volatile unsigned *)g_dbg_word = code; __threadfence_system(); } asm volatile("trap;"); }
}
__device__ __forceinline__ unsigned ll_wait(const LLf *p, unsigned tag, unsigned code) {
    unsigned long long w; unsigned spins = 0;
    for (;;) { w = ll_load(p); if ((unsigned)(w >> 32) == tag) break; spin_guard(spins, code); __nanosleep(40); }
    return (unsigned)w;
}
// hint: sleep-poll ONE word until `target` CTAs have finished the producing op, then look at the tagged data (normally ready at the first look).
// Polling the data itself from 148 x 256 threads saturates the L2 request path and slows the weight stream (measured: 272 -> 221 us per token
// just by not polling while the CTA's own warps still consume).
__device__ __forceinline__ void hint_wait(const unsigned *counter, unsigned target, unsigned code) {
    unsigned v, spins = 0;
    for (;;) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        if ((int)(v - target) >= 0) break;
        spin_guard(spins, code); __nanosleep(60);
    }
}
__device__ __forceinline__ void hint_post(unsigned *counter) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory"); }
// two elements with one 16-byte access (each 8-byte half was written by one 64-bit store, so it is seen whole)
__device__ __forceinline__ void ll_load2(const LLf *p, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ bool ll_try4(const LLf *p, unsigned tag, float4 &out) {
    unsigned long long a, b, c, d;
    ll_load2(p, a, b); ll_load2(p + 2, c, d);
    out = make_float4(__uint_as_float((unsigned)a), __uint_as_float((unsigned)b), __uint_as_float((unsigned)c), __uint_as_float((unsigned)d));
    return (unsigned)(a >> 32) == tag && (unsigned)(b >> 32) == tag && (unsigned)(c >> 32) == tag && (unsigned)(d >> 32) == tag;
}

// quantise the float4 at elements i..i+3 (8 consecutive lanes cover one 32-element block); identical bytes to k::stage_act
template <int ACT>
__device__ __forceinline__ void quant_item5(const float4 a, int i, bool valid, int cols, unsigned char *sm) {
    float *d = (float *)(sm + act5_d(cols)); float *s = d + cols / 32;
    const int j8 = threadIdx.x & 7, b = i >> 5;
    float amax = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4)); amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2)); amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    const float dd = amax / 127.f;
    const float id = amax != 0.0f ? 127.f / amax : 0.0f;
    const int q0 = __float2int_rn(a.x * id), q1 = __float2int_rn(a.y * id), q2 = __float2int_rn(a.z * id), q3 = __float2int_rn(a.w * id);
    int sum = (q0 + q1) + (q2 + q3);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4); sum += __shfl_xor_sync(0xffffffffu, sum, 2); sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    if (valid) {
        *(unsigned *)(sm + (j8 < 4 ? 0 : act5_hi(cols)) + b * 16 + (j8 & 3) * 4) = (unsigned)(q0 & 0xff) | ((unsigned)(q1 & 0xff) << 8) | ((unsigned)(q2 & 0xff) << 16) | ((unsigned)(q3 & 0xff) << 24);
        if (j8 == 0) {
            if (ACT == ACT_Q8_0) { d[b] = __half2float(__float2half_rn(dd)); s[b] = (float)sum; }  // integer block sum (exact): Q4_0's "-8" term
            else { d[b] = dd; s[b] = dd * (float)sum; }
        }
    }
}

// un-normed n_ff-wide input (down <- act), all 480 consumer threads: thread t owns the float4s 480 k + t
template <int ACT>
__device__ __forceinline__ void stage_plain5(const LLf *__restrict__ x, unsigned tag, int cols, unsigned char *sm, long long *tr) {
    const int tid = threadIdx.x;
    float4 xv[kPlainItems];
    unsigned need = 0;
#pragma unroll
    for (int it = 0; it < kPlainItems; ++it) { xv[it] = make_float4(0.f, 0.f, 0.f, 0.f); if (4 * (tid + kConsumerThreads * it) < cols) need |= 1u << it; }
    unsigned spins = 0;
    while (need) {
        spin_guard(spins, 0x100u);
#pragma unroll
        for (int it = 0; it < kPlainItems; ++it)
            if (need & (1u << it)) { if (ll_try4(x + 4 * (tid + kConsumerThreads * it), tag, xv[it])) need &= ~(1u << it); }
    }
    if (tr) tr[1] = clock64();
#pragma unroll
    for (int it = 0; it < kPlainItems; ++it) {
        const int i = 4 * (tid + kConsumerThreads * it);
        if (4 * kConsumerThreads * it < cols) quant_item5<ACT>(xv[it], i, i < cols, cols, sm);  // (CTA-uniform predicate)
    }
}
// RMS-normed n_embd-wide input (qkv, gate/up, output <- x), warps 0-7; element ownership and reduction order of k::stage_act
template <int ACT>
__device__ __forceinline__ void stage_norm5(const LLf *__restrict__ x, unsigned tag, const float *__restrict__ nw, int cols, unsigned char *sm, double *red, long long *tr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;  // tid < 256
    float4 xv[kNormItems];
    unsigned need = 0;
#pragma unroll
    for (int it = 0; it < kNormItems; ++it) { xv[it] = make_float4(0.f, 0.f, 0.f, 0.f); if (1024 * it + 4 * tid < cols) need |= 1u << it; }
    unsigned spins = 0;
    while (need) {
        spin_guard(spins, 0x200u);
#pragma unroll
        for (int it = 0; it < kNormItems; ++it)
            if (need & (1u << it)) { if (ll_try4(x + 1024 * it + 4 * tid, tag, xv[it])) need &= ~(1u << it); }
    }
    if (tr) tr[1] = clock64();
    float4 w[kNormItems];
    mk::load_norm_weights(nw, cols, w);
    double ssa = 0.0, ssb = 0.0;
#pragma unroll
    for (int it = 0; it < kNormItems; ++it) {
        if (1024 * it + 4 * tid < cols) {
            const float4 a = xv[it];
            if (it & 1) { ssb += (double)(a.x * a.x); ssb += (double)(a.y * a.y); ssb += (double)(a.z * a.z); ssb += (double)(a.w * a.w); }
            else        { ssa += (double)(a.x * a.x); ssa += (double)(a.y * a.y); ssa += (double)(a.z * a.z); ssa += (double)(a.w * a.w); }
        }
    }
    ssa = warp_sum(ssa); ssb = warp_sum(ssb);
    if (lane == 0) { red[warp] = ssa; red[warp + 8] = ssb; }
    cta_sync<true>();
    if (warp == 0) { double t = lane < 16 ? red[lane] : 0.0; t = warp_sum(t); if (lane == 0) red[32] = t; }
    cta_sync<true>();
    const double tot = red[32];
    const float mean = (float)(tot / (double)cols);
    const float scale = 1.0f / sqrtf(mean + 1e-6f);
#pragma unroll
    for (int it = 0; it < kNormItems; ++it) {
        const int i = 1024 * it + 4 * tid;
        if (1024 * it < cols) {  // (CTA-uniform predicate)
            const float4 a = xv[it], w4 = w[it];
            quant_item5<ACT>(make_float4((a.x * scale) * w4.x, (a.y * scale) * w4.y, (a.z * scale) * w4.z, (a.w * scale) * w4.w), i, i < cols, cols, sm);
        }
    }
}
// wo <- the quantised attention output: nb*8 words of 4 int8, nb d, nb s, all tagged (written by attention5); warps 0-7
constexpr int kAttItems = 7;  // 256 threads x 7 >= 160 blocks x 10 words (n_embd <= 5120)
__device__ __forceinline__ void stage_att5(const LLf *__restrict__ att, unsigned tag, int cols, unsigned char *sm, long long *tr) {
    const int tid = threadIdx.x, nb = cols >> 5, total = nb * 10;  // tid < 256
    unsigned need = 0, val[kAttItems];
#pragma unroll
    for (int it = 0; it < kAttItems; ++it) { val[it] = 0u; if (tid + 256 * it < total) need |= 1u << it; }
    unsigned spins = 0;
    while (need) {
        spin_guard(spins, 0x300u);
#pragma unroll
        for (int it = 0; it < kAttItems; ++it)
            if (need & (1u << it)) { const unsigned long long w = ll_load(att + tid + 256 * it); if ((unsigned)(w >> 32) == tag) { val[it] = (unsigned)w; need &= ~(1u << it); } }
    }
    if (tr) tr[1] = clock64();
    float *d = (float *)(sm + act5_d(cols));
#pragma unroll
    for (int it = 0; it < kAttItems; ++it) {
        const int w = tid + 256 * it;
        if (w < nb * 8) { const int B = w >> 3, j = w & 7; *(unsigned *)(sm + (j < 4 ? 0 : act5_hi(cols)) + B * 16 + (j & 3) * 4) = val[it]; }
        else if (w < total) d[w - nb * 8] = __uint_as_float(val[it]);  // d[0..nb) then s[0..nb) are contiguous
    }
}

// ---- dot products -----------------------------------------------------------------------------------------------------------------
// shared-memory activations (n_ff-wide inputs, and every input of models whose n_embd is not 1024 * NBL): k::dot2_q4 order
template <bool Q41>
__device__ __forceinline__ void dot2_q4_smem(const unsigned char *row0, const unsigned char *row1, int nb, int cols, const unsigned char *act, int lane, float &r0, float &r1) {
    const uint4 *qs0 = (const uint4 *)row0, *qs1 = (const uint4 *)row1;
    const unsigned char *sc0 = row0 + (size_t)nb * 16, *sc1 = row1 + (size_t)nb * 16;
    const int4 *alo = (const int4 *)act, *ahi = (const int4 *)(act + act5_hi(cols));
    const float *ad = (const float *)(act + act5_d(cols)), *as = ad + nb;
    float accd0 = 0.f, accd1 = 0.f, accm0 = 0.f, accm1 = 0.f;
#pragma unroll 1
    for (int b = lane; b < nb; b += 32) {
        const uint4 q0 = qs0[b], q1 = qs1[b];
        const int4 la = alo[b], ha = ahi[b];
        const float adv = ad[b], asv = as[b];
        int s0 = q4_block_idot(q0, la, ha), s1 = q4_block_idot(q1, la, ha);
        if (Q41) {
            const float2 f0 = __half22float2(((const __half2 *)sc0)[b]), f1 = __half22float2(((const __half2 *)sc1)[b]);
            accd0 = fmaf(f0.x * adv, (float)s0, accd0); accm0 = fmaf(f0.y, asv, accm0);
            accd1 = fmaf(f1.x * adv, (float)s1, accd1); accm1 = fmaf(f1.y, asv, accm1);
        } else {
            const float d0 = __half2float(((const __half *)sc0)[b]), d1 = __half2float(((const __half *)sc1)[b]);
            const int i8 = 8 * (int)asv;
            s0 -= i8; s1 -= i8;
            accd0 += ((float)s0 * d0) * adv; accd1 += ((float)s1 * d1) * adv;
        }
    }
    r0 = warp_sum(accd0) + warp_sum(accm0);
    r1 = warp_sum(accd1) + warp_sum(accm1);
}
// one row (the n_ff-wide `down` matrix: a unit is one row, three rows share a ring slot)
template <bool Q41>
__device__ __forceinline__ float dot1_q4_smem(const unsigned char *row0, int nb, int cols, const unsigned char *act, int lane) {
    const uint4 *qs0 = (const uint4 *)row0;
    const unsigned char *sc0 = row0 + (size_t)nb * 16;
    const int4 *alo = (const int4 *)act, *ahi = (const int4 *)(act + act5_hi(cols));
    const float *ad = (const float *)(act + act5_d(cols)), *as = ad + nb;
    float accd0 = 0.f, accm0 = 0.f;
#pragma unroll 4
    for (int b = lane; b < nb; b += 32) {
        const uint4 q0 = qs0[b];
        const int4 la = alo[b], ha = ahi[b];
        const float adv = ad[b], asv = as[b];
        int s0 = q4_block_idot(q0, la, ha);
        if (Q41) {
            const float2 f0 = __half22float2(((const __half2 *)sc0)[b]);
            accd0 = fmaf(f0.x * adv, (float)s0, accd0); accm0 = fmaf(f0.y, asv, accm0);
        } else {
            const float d0 = __half2float(((const __half *)sc0)[b]);
            s0 -= 8 * (int)asv;
            accd0 += ((float)s0 * d0) * adv;
        }
    }
    return warp_sum(accd0) + warp_sum(accm0);
}
// register-resident activations: lane l holds blocks l + 32 i (i < NBL) of the staged vector; same per-lane order and arithmetic
template <int NBL> struct ActRegs { int4 lo[NBL], hi[NBL]; float d[NBL], s[NBL]; };
template <int NBL>
__device__ __forceinline__ void load_act_regs(const unsigned char *act, int cols, int lane, ActRegs<NBL> &r) {
    const int4 *alo = (const int4 *)act, *ahi = (const int4 *)(act + act5_hi(cols));
    const float *ad = (const float *)(act + act5_d(cols)), *as = ad + (cols >> 5);
#pragma unroll
    for (int i = 0; i < NBL; ++i) { const int b = lane + 32 * i; r.lo[i] = alo[b]; r.hi[i] = ahi[b]; r.d[i] = ad[b]; r.s[i] = as[b]; }
}
template <bool Q41, int NBL>
__device__ __forceinline__ void dot2_q4_reg(const unsigned char *row0, const unsigned char *row1, const ActRegs<NBL> &a, int lane, float &r0, float &r1) {
    constexpr int nb = 32 * NBL;
    const uint4 *qs0 = (const uint4 *)row0, *qs1 = (const uint4 *)row1;
    const unsigned char *sc0 = row0 + (size_t)nb * 16, *sc1 = row1 + (size_t)nb * 16;
    float accd0 = 0.f, accd1 = 0.f, accm0 = 0.f, accm1 = 0.f;
#pragma unroll
    for (int i = 0; i < NBL; ++i) {
        const int b = lane + 32 * i;
        const uint4 q0 = qs0[b], q1 = qs1[b];
        int s0 = q4_block_idot(q0, a.lo[i], a.hi[i]), s1 = q4_block_idot(q1, a.lo[i], a.hi[i]);
        if (Q41) {
            const float2 f0 = __half22float2(((const __half2 *)sc0)[b]), f1 = __half22float2(((const __half2 *)sc1)[b]);
            accd0 = fmaf(f0.x * a.d[i], (float)s0, accd0); accm0 = fmaf(f0.y, a.s[i], accm0);
            accd1 = fmaf(f1.x * a.d[i], (float)s1, accd1); accm1 = fmaf(f1.y, a.s[i], accm1);
        } else {
            const float d0 = __half2float(((const __half *)sc0)[b]), d1 = __half2float(((const __half *)sc1)[b]);
            const int i8 = 8 * (int)a.s[i];
            s0 -= i8; s1 -= i8;
            accd0 += ((float)s0 * d0) * a.d[i]; accd1 += ((float)s1 * d1) * a.d[i];
        }
    }
    r0 = warp_sum(accd0) + warp_sum(accm0);
    r1 = warp_sum(accd1) + warp_sum(accm1);
}

// unit -> CTA mapping: CTA c owns the contiguous units [c n / G, (c + 1) n / G), so that a group of consecutive units is one contiguous copy
__device__ __forceinline__ int cta_units(int cta, int n_su, int G) { return unit_begin(cta + 1, n_su, G) - unit_begin(cta, n_su, G); }

// shared memory carve-up (dynamic)
struct Smem5 { unsigned char *ring, *ff, *e0; uint64_t *full, *empty; Op5 *ops; };  // full: [2][n_slots] (even / odd ring laps), empty: [n_slots]
__device__ __forceinline__ Smem5 carve5(const Params &P) {
    extern __shared__ __align__(128) unsigned char smem[];
    Smem5 m;
    m.ring = smem; m.ff = smem + (size_t)P.n_slots * P.slot_bytes;
    m.e0 = m.ff + P.ff_bytes;
    m.full = (uint64_t *)(m.e0 + P.e_bytes); m.empty = m.full + 2 * P.n_slots;
    m.ops = (Op5 *)(m.empty + P.n_slots);
    return m;
}

// The matvec phase of one op for one consumer warp: outputs are tagged.  Unit k of the CTA's share sits in fill (n_base + k / upg) at offset
// (k % upg) * unit bytes; warp w takes units w, w + W, ...
// Every slot has TWO "full" barriers, used by even and odd ring laps.  With one, a warp that waits for fill n could see the completed phase of
// fill n - 2 laps while fill n - 1 lap is still in flight (same parity: the async copies of different fills may complete out of order under
// load), read a half-written slot and release it.  With two, the previous phase of fill n's barrier is fill n - 2 laps, which must have been
// consumed before fill n - 1 lap could be issued, and that fill was issued before the fill of this warp's previous unit (fills are issued in
// order; the previous unit is less than a lap back, or in the previous op, all of whose fills are consumed).
template <bool Q41, int KIND, int NBL, bool TRACE>
__device__ __forceinline__ unsigned consume5(const Params &P, int oi, unsigned n_base, int pos, unsigned tag, const unsigned char *actb, long long *tr) {
    const Smem5 m = carve5(P);
    const Op5 &op = m.ops[oi];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int W = op.n_warps, upg = op.upg;
    const int lo = unit_begin(cta, op.n_su, G), cnt = unit_begin(cta + 1, op.n_su, G) - lo;
    const unsigned n_next = n_base + (unsigned)((cnt + upg - 1) / upg);
    if (warp >= W) return n_next;
    const int cols = op.cols, nb = cols >> 5, S = P.n_slots;
    constexpr bool REG = NBL > 0 && KIND != OP_DOWN;  // n_embd-wide input held in registers
    constexpr int RPU = KIND == OP_DOWN ? 1 : 2;       // rows per unit
    ActRegs<REG ? NBL : 1> ar;
    if (REG) load_act_regs<REG ? NBL : 1>(actb, cols, lane, ar);
    const unsigned rb = (unsigned)op.row_bytes, slot_bytes = (unsigned)P.slot_bytes;
    unsigned long long best = 0ull;
    __half pend_h = __ushort_as_half((unsigned short)0); float pend_up = 0.f; int pend_i = -1;
    long long t_wait = 0, t_dot = 0, t_epi = 0; int n_units = 0;  // (TRACE: kept in registers, stored once per op)
    for (int k = warp; k < cnt; k += W) {
        const int g = k / upg, sub = k - g * upg;
        const unsigned n = n_base + (unsigned)g;
        const int sl = (int)(n % (unsigned)S);
        const unsigned lap = n / (unsigned)S;
        uint64_t *const fb = &m.full[(lap & 1u) * (unsigned)S + (unsigned)sl];
        const unsigned ph = (lap >> 1) & 1u;
        const int r0 = (lo + k) * RPU;
        float2 rs = make_float2(0.f, 0.f);
        if (KIND == OP_WO) {  // residual rows: final since this CTA gathered the whole of x for the previous normed op
            const uint4 t = __ldcg((const uint4 *)(P.x + r0));
            rs = make_float2(__uint_as_float(t.x), __uint_as_float(t.z));
        }
        if (KIND == OP_DOWN) rs.x = __uint_as_float(__ldcg((const unsigned *)(P.x + r0)));
        if (KIND == OP_QKV) { if (r0 < 2 * P.E) rs = __ldg(&P.rope[(size_t)pos * 64 + ((r0 % P.E) % 128) / 2]); }
        long long tw0 = 0, tw1 = 0, tw2 = 0;
        if (TRACE && tr) tw0 = clock64();
        mb_wait(fb, ph);
        if (TRACE && tr) tw1 = clock64();
        const unsigned char *row0 = m.ring + (size_t)sl * slot_bytes + (size_t)sub * (RPU * rb);
        const unsigned char *row1 = row0 + rb;
        float v0, v1 = 0.f;
        if (KIND == OP_DOWN) v0 = dot1_q4_smem<Q41>(row0, nb, cols, actb, lane);
        else if (REG) dot2_q4_reg<Q41, REG ? NBL : 1>(row0, row1, ar, lane, v0, v1);
        else dot2_q4_smem<Q41>(row0, row1, nb, cols, actb, lane, v0, v1);
        if (TRACE && tr) { tw2 = clock64(); t_wait += tw1 - tw0; t_dot += tw2 - tw1; ++n_units; }
        if (lane == 0) {
            mb_arrive(&m.empty[sl]);
            if (KIND == OP_QKV) {
                const int E = P.E, partn = r0 / E, rr = r0 % E;
                const size_t kvo = ((size_t)op.layer * P.n_ctx + pos) * E + rr;
                if (partn == 2) {
                    const __half2 h2 = __floats2half2_rn(v0, v1);
                    *(__half2 *)(P.vcache + kvo) = h2;                                  // for later tokens (stream order)
                    ll_store_bits(P.vcur + (rr >> 1), *(const unsigned *)&h2, tag);     // for this token's attention
                } else {
                    const float2 cs = rs;
                    const float o0 = v0 * cs.x - v1 * cs.y, o1 = v0 * cs.y + v1 * cs.x;
                    if (partn == 0) { ll_store(P.q + rr, o0, tag); ll_store(P.q + rr + 1, o1, tag); }
                    else {
                        const __half2 h2 = __floats2half2_rn(o0, o1);
                        *(__half2 *)(P.kcache + kvo) = h2;
                        ll_store_bits(P.kcur + (rr >> 1), *(const unsigned *)&h2, tag);
                    }
                }
            } else if (KIND == OP_WO) {
                ll_store(P.x + r0, v0 + rs.x, tag); ll_store(P.x + r0 + 1, v1 + rs.y, tag);
            } else if (KIND == OP_DOWN) {
                ll_store(P.x + r0, v0 + rs.x, tag);
            } else if (KIND == OP_GATEUP) {
                if (pend_i >= 0) ll_store(P.act + pend_i, __half2float(pend_h) * pend_up, tag);
                pend_h = P.tab_silu[__half_as_ushort(__float2half_rn(v0))]; pend_up = v1; pend_i = r0 >> 1;
            } else {  // OP_OUTPUT
                P.logits[r0] = v0;
                const unsigned long long k0 = argmax_key(v0, r0);
                best = best > k0 ? best : k0;
                if (r0 + 1 < P.n_vocab) { P.logits[r0 + 1] = v1; const unsigned long long k1 = argmax_key(v1, r0 + 1); best = best > k1 ? best : k1; }
            }
        }
        if (TRACE && tr) t_epi += clock64() - tw2;
    }
    if (KIND == OP_GATEUP) { if (lane == 0 && pend_i >= 0) ll_store(P.act + pend_i, __half2float(pend_h) * pend_up, tag); }
    if (KIND == OP_OUTPUT) { if (lane == 0 && best) atomicMax(&P.state->argmax_key, best); }
    if (TRACE && tr) { tr[4] = t_wait; tr[5] = t_dot; tr[6] = n_units; tr[7] = t_epi; }
    return n_next;
}

// ---- attention of one head (256 threads = warps 0-7 of CTA h), k::attention_head with three changes: q and the current position's
// K / V arrive as tagged words (the cache row of `pos` may not be visible yet: no barrier has been crossed since it was written), and
// the 128 outputs leave as the four Q8 blocks the `wo` matvec consumes.  All float orders are those of k::attention_head.
template <int ACT>
__device__ __noinline__ void attention5(const Params &P, int layer, int pos, int h, unsigned tag_in, unsigned tag_out, unsigned char *dyn,
                                         double *red, float *redf, float *qs, __half *kcur_s, __half *vcur_s, float *part, long long *tr) {
    constexpr int B = 12;
    const int E = P.E, n_ctx = P.n_ctx;
    const __half *kc = P.kcache + (size_t)layer * n_ctx * E, *vc = P.vcache + (size_t)layer * n_ctx * E;
    float *sc = (float *)dyn; __half *ph = (__half *)(dyn + (size_t)n_ctx * 4);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nkv = pos + 1;
    const int sub = lane >> 4, l16 = lane & 15;
    uint4 kv[B];
    if (tid < 128) qs[tid] = __uint_as_float(ll_wait(P.q + h * 128 + tid, tag_in, 0x400u));
    else if (tid < 192) ((unsigned *)kcur_s)[tid - 128] = ll_wait(P.kcur + h * 64 + (tid - 128), tag_in, 0x401u);
    else ((unsigned *)vcur_s)[tid - 192] = ll_wait(P.vcur + h * 64 + (tid - 192), tag_in, 0x402u);
    cta_sync<true>();
    if (tr) tr[1] = clock64();
    {
        __half2 q2[4];
        {
            const float4 qa = *(const float4 *)(qs + l16 * 8), qb = *(const float4 *)(qs + l16 * 8 + 4);
            q2[0] = __floats2half2_rn(qa.x, qa.y); q2[1] = __floats2half2_rn(qa.z, qa.w); q2[2] = __floats2half2_rn(qb.x, qb.y); q2[3] = __floats2half2_rn(qb.z, qb.w);
        }
        const uint4 kcur4 = *(const uint4 *)(kcur_s + l16 * 8);
        for (int kb0 = warp * 2; kb0 < nkv; kb0 += 16 * B) {  // warp-uniform trip counts (both half-warps shuffle together)
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const int key = min(kb0 + u * 16 + sub, nkv - 1);  // clamped, unconditional (a predicated load would demote kv[] to local memory)
                kv[u] = ld_kv16<true>(kc + (size_t)key * E + h * 128 + l16 * 8);
            }
#pragma unroll
            for (int u = 0; u < B; ++u) {
                if (kb0 + u * 16 >= nkv) break;
                const int key = kb0 + u * 16 + sub;
                float s = 0.f;
                if (key < nkv) {
                    const uint4 kk = key == pos ? kcur4 : kv[u];
                    const __half2 *k2 = (const __half2 *)&kk;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float2 a = __half22float2(k2[j]), b = __half22float2(q2[j]); s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); }
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (key < nkv && l16 == 0) sc[key] = s * P.kq_scale;
            }
        }
    }
    // the first batch of V rows does not depend on the scores: request it now (runs under the soft-max)
    constexpr int BV = 12;
    const uint4 vcur4 = *(const uint4 *)(vcur_s + (tid & 15) * 8);
    uint4 vv0[BV];
#pragma unroll
    for (int u = 0; u < BV; ++u) { const int key = min((tid >> 4) + 16 * u, nkv - 1); vv0[u] = ld_kv16<true>(vc + (size_t)key * E + h * 128 + (tid & 15) * 8); }
    cta_sync<true>();
    float mx = -INFINITY;
    for (int i = tid; i < nkv; i += 256) mx = fmaxf(mx, sc[i]);
    mx = block_max<true>(mx, redf);
    double sum = 0.0;
    for (int i = tid; i < nkv; i += 256) { const float v = lut_f16(P.tab_exp, sc[i] - mx); sc[i] = v; sum += (double)v; }
    const double tot = block_sum<true>(sum, red);
    const float inv = (float)(1.0 / tot);
    for (int i = tid; i < nkv; i += 256) ph[i] = __float2half_rn(sc[i] * inv);
    cta_sync<true>();
    {   // P.V : thread = (key group g of 16, dim octet o of 16); groups are combined by a pairwise tree (canonical order)
        const int g = tid >> 4, o = tid & 15;
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
        for (int u = 0; u < BV; ++u) {
            const int key = g + 16 * u;
            if (key < nkv) {
                const float p = __half2float(ph[key]);
                const uint4 vk = key == pos ? vcur4 : vv0[u];
                const __half2 *v2 = (const __half2 *)&vk;
#pragma unroll
                for (int j = 0; j < 4; ++j) { const float2 v = __half22float2(v2[j]); acc[2 * j] = fmaf(v.x, p, acc[2 * j]); acc[2 * j + 1] = fmaf(v.y, p, acc[2 * j + 1]); }
            }
        }
        for (int key0 = g + 16 * BV; key0 < nkv; key0 += 16 * B) {
            uint4 vv[B];
#pragma unroll
            for (int u = 0; u < B; ++u) { const int key = min(key0 + 16 * u, nkv - 1); vv[u] = ld_kv16<true>(vc + (size_t)key * E + h * 128 + o * 8); }
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const int key = key0 + 16 * u;
                if (key < nkv) {
                    const float p = __half2float(ph[key]);
                    const uint4 vk = key == pos ? vcur4 : vv[u];
                    const __half2 *v2 = (const __half2 *)&vk;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { const float2 v = __half22float2(v2[j]); acc[2 * j] = fmaf(v.x, p, acc[2 * j]); acc[2 * j + 1] = fmaf(v.y, p, acc[2 * j + 1]); }
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) part[g * 128 + o * 8 + e] = acc[e];
    }
    cta_sync<true>();
    if (tid < 128) {  // warp w = dims 32 w .. 32 w + 31 = Q8 block w of this head
        float v[16];
#pragma unroll
        for (int g = 0; g < 16; ++g) v[g] = part[g * 128 + tid];
#pragma unroll
        for (int st = 1; st < 16; st <<= 1)
#pragma unroll
            for (int g = 0; g < 16; g += 2 * st) v[g] = v[g] + v[g + st];
        const float val = v[0];
        const float amax = warp_max(fabsf(val));
        const float dd = amax / 127.f;
        const float id = amax != 0.0f ? 127.f / amax : 0.0f;
        const int qv = __float2int_rn(val * id);
        const int isum = warp_sum(qv);
        const unsigned byte = (unsigned)(qv & 0xff);
        const unsigned b1 = __shfl_down_sync(0xffffffffu, byte, 1), b2 = __shfl_down_sync(0xffffffffu, byte, 2), b3 = __shfl_down_sync(0xffffffffu, byte, 3);
        const int Bk = h * 4 + warp, nbE = E >> 5;
        if ((lane & 3) == 0) ll_store_bits(P.att + Bk * 8 + (lane >> 2), byte | (b1 << 8) | (b2 << 16) | (b3 << 24), tag_out);
        if (lane == 0) {
            if (ACT == ACT_Q8_0) { ll_store(P.att + nbE * 8 + Bk, __half2float(__float2half_rn(dd)), tag_out); ll_store(P.att + nbE * 9 + Bk, (float)isum, tag_out); }
            else { ll_store(P.att + nbE * 8 + Bk, dd, tag_out); ll_store(P.att + nbE * 9 + Bk, dd * (float)isum, tag_out); }
        }
    }
}

// The producer thread: for every op with weights, the CTA's share of rows is one contiguous range; it is cut into groups of upg units and
// every group is ONE bulk copy into the next ring slot.  The slot's "empty" barrier expects upg_max arrivals: the units of the group arrive as
// they are consumed, the producer adds the difference for smaller groups.  Running pointers only: everything stays in registers.
__device__ __forceinline__ void mb_arrive_n(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory"); }
__device__ __noinline__ void producer5(const Params &P) {
    const Smem5 m = carve5(P);
    if ((threadIdx.x & 31) != 0) return;
    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    unsigned s = 0, ph = 0, lap = 0;
    long long t_blocked = 0, t_begin = 0; unsigned n_chunks = 0;
    const bool stats = P.trace != nullptr && (cta == 0 || cta == G - 1);
    if (stats) t_begin = clock64();
    for (int oi = 0; oi < P.n_ops; ++oi) {
        const unsigned char *w = m.ops[oi].w;
        if (!w) continue;
        const int n_su = m.ops[oi].n_su, upg = m.ops[oi].upg;
        const unsigned ub = (unsigned)m.ops[oi].row_bytes * (unsigned)m.ops[oi].rpu;  // bytes per unit
        const int lo = unit_begin(cta, n_su, G), hi = unit_begin(cta + 1, n_su, G);
        const unsigned char *src = w + (size_t)lo * ub;
        for (int left = hi - lo; left > 0; left -= upg) {
            const int nu = left < upg ? left : upg;
            const unsigned bytes = (unsigned)nu * ub;
            long long t0 = 0;
            if (stats) t0 = clock64();
            mb_wait(&m.empty[s], ph ^ 1u);
            if (stats) { t_blocked += clock64() - t0; ++n_chunks; }
            uint64_t *const fb = &m.full[(lap & 1u) * (unsigned)P.n_slots + s];
            mb_expect_tx(fb, bytes);
            bulk_g2s(m.ring + (size_t)s * P.slot_bytes, src, bytes, fb);
            if (nu < P.upg_max) mb_arrive_n(&m.empty[s], (uint32_t)(P.upg_max - nu));
            src += bytes;
            if (++s == (unsigned)P.n_slots) { s = 0; ph ^= 1u; ++lap; }
        }
    }
    if (stats) {  // after the per-op records: [2 CTAs][8]: cycles waiting for a free slot, total cycles, copies, unused
        long long *o = P.trace + (size_t)2 * P.n_ops * 8 + (cta == 0 ? 0 : 8);
        o[0] = t_blocked; o[1] = clock64() - t_begin; o[2] = (long long)n_chunks; o[3] = 0;
    }
}

// ---- the kernel --------------------------------------------------------------------------------------------------------------------
// Why no write-after-read hazard appears without grid barriers: every consumer of a vector gathers ALL of its elements, and every producer
// of the next version of a vector transitively depends on such a full gather by EVERY CTA (x' needs all of att; att needs all heads' q/k/v,
// whose rows are spread over all CTAs; each of those CTAs gathered all of x before producing a row; act needs all of x'; x'' needs all of
// act, ...), so all reads of version n have completed on every CTA before any CTA can produce an element of version n+1.  Tags are unique per
// (launch, op), hence a stale element can only look "not ready", never "ready".  Shared-memory reuse: every op has exactly one CTA-wide barrier
// (after staging), which a warp reaches only after it finished the previous op, and no staging buffer is reused by consecutive ops.
template <int WT, int NBL, bool TRACE>
__global__ void __launch_bounds__(kMegaThreads, 1) decode_megakernel5(const __grid_constant__ Params P) {
    __shared__ double red[34];
    __shared__ float redf[34];
    __shared__ __align__(16) float qs[128];
    __shared__ __align__(16) __half kcur_s[128];
    __shared__ __align__(16) __half vcur_s[128];
    __shared__ float part[16 * 128];
    __shared__ unsigned op_done;   // consumer warps that finished their units, all ops so far (the 15th of an op posts the CTA's completion hint)
    constexpr int ACT = act_of(WT);
    constexpr bool Q41 = WT == GG_Q4_1;
    const Smem5 m = carve5(P);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int G = (int)gridDim.x, cta = (int)blockIdx.x;
    if (tid == 0) { op_done = 0u; g_dbg_word = (unsigned *)P.dbg; }

    if (tid == 0) {
        for (int s = 0; s < P.n_slots; ++s) { mb_init(&m.full[s], 1); mb_init(&m.full[P.n_slots + s], 1); mb_init(&m.empty[s], (uint32_t)P.upg_max); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < P.n_ops * 2; i += kMegaThreads) ((uint4 *)m.ops)[i] = ((const uint4 *)P.ops)[i];
    __syncthreads();  // the only CTA-wide barrier; afterwards consumers use named barriers 2 (480 threads) and 1 (256 threads)

    if (warp == kConsumerWarps) { producer5(P); return; }

    unsigned n_base = 0;
    const int pos = __ldcg(&P.state->n_past);
    const unsigned seqv = __ldcg(P.seq);
    const unsigned tag0 = (seqv << 10) + 1u;  // tag of op oi in this launch = tag0 + oi
    const unsigned all_ctas = (seqv + 1u) * (unsigned)G, all_heads = (seqv + 1u) * (unsigned)P.n_head;  // hint targets (counters are monotonic)
    unsigned tag_x = 0, tag_att = 0, tag_act = 0, tag_qkv = 0;  // tag of the op that last produced each vector
    for (int oi = 0; oi < P.n_ops; ++oi) {
        const int kind = m.ops[oi].kind;
        const unsigned tag = tag0 + (unsigned)oi;
        long long *tr = nullptr;
        if (TRACE) { if (P.trace && tid == 0 && (cta == 0 || cta == G - 1)) tr = P.trace + ((size_t)(cta == 0 ? 0 : 1) * P.n_ops + oi) * 8; }
        if (TRACE && tr) { tr[0] = clock64(); tr[1] = tr[0]; tr[2] = 0; tr[3] = 0; tr[4] = 0; tr[5] = 0; tr[6] = 0; tr[7] = 0; }
        if (kind == OP_EMBED) {
            const int token = __ldcg(&P.state->tokens[0]);
            const unsigned char *row = P.tok + (size_t)token * P.tok_row_bytes;
            for (int i = cta * kConsumerThreads + tid; i < P.E; i += G * kConsumerThreads) ll_store(P.x + i, dequant_elem(P.tok_type, row, i), tag);
            if (tid == 0) hint_post(P.done + oi);
            tag_x = tag;
            continue;
        }
        if (kind == OP_ATTN) {
            if (cta < P.n_head && tid < 256) {
                hint_wait(P.done + (tag_qkv - tag0), all_ctas, 0x500u);
                attention5<ACT>(P, m.ops[oi].layer, pos, cta, tag_qkv, tag, m.ff, red, redf, qs, kcur_s, vcur_s, part, TRACE ? tr : nullptr);
                if (tid == 0) hint_post(P.done + oi);
            }
            tag_att = tag;
            continue;
        }
        if (kind == OP_FINAL) {
            grid_barrier(P.barrier, (unsigned)G);  // the only grid barrier of the launch: every CTA's logits / arg-max candidates are in
            if (cta == 0 && tid == 0) {
                DeviceState *st = P.state;
                const unsigned long long key = __ldcg((const unsigned long long *)&st->argmax_key);
                const int id = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
                st->argmax_id = id; st->tokens[0] = id; st->argmax_key = 0ull;
                st->n_past += 1; st->n_tok = 1;
                *P.seq = (tag0 >> 10) + 1u;
            }
            continue;
        }
        // ---- matvec ops: [gather + stage the input] -> one CTA barrier -> [ring slots -> dot -> tagged epilogue] ----
        const Op5 &op = m.ops[oi];
        const int cols = op.cols;
        unsigned char *actb;
        if (P.flags & 4) consumer_sync();
        if (kind == OP_DOWN) {
            actb = m.ff;
            hint_wait(P.done + (tag_act - tag0), all_ctas, 0x501u);
            stage_plain5<ACT>(P.act, tag_act, cols, actb, TRACE ? tr : nullptr);
        } else {
            // qkv / gate-up / output stage into e0, wo into the ff region (the attention scratch is dead by then; `down` restages it two
            // CTA barriers later): between two uses of a buffer lies a barrier that every warp reaches only after it is done reading
            actb = kind == OP_WO ? m.ff : m.e0;
            if (tid < 256) {
                if (kind == OP_WO) { hint_wait(P.done + (tag_att - tag0), all_heads, 0x502u); stage_att5(P.att, tag_att, cols, actb, TRACE ? tr : nullptr); }
                else {
                    if (tid * 32 < cols) prefetch_l2(op.norm_w + tid * 32);
                    hint_wait(P.done + (tag_x - tag0), all_ctas, 0x503u);
                    stage_norm5<ACT>(P.x, tag_x, op.norm_w, cols, actb, red, TRACE ? tr : nullptr);
                }
            }
        }
        consumer_sync();
        if (TRACE && tr) tr[2] = clock64();
        if (kind == OP_QKV && (P.flags & 1) && cta < P.n_head && tid < 256) {
            const size_t lo = (size_t)op.layer * P.n_ctx * P.E;
            prefetch_kv_head(P.kcache + lo, P.vcache + lo, pos, cta, P.E);
        }
        switch (kind) {
            case OP_QKV:    n_base = consume5<Q41, OP_QKV, NBL, TRACE>(P, oi, n_base, pos, tag, actb, tr); tag_qkv = tag; break;
            case OP_WO:     n_base = consume5<Q41, OP_WO, NBL, TRACE>(P, oi, n_base, pos, tag, actb, tr); tag_x = tag; break;
            case OP_GATEUP: n_base = consume5<Q41, OP_GATEUP, NBL, TRACE>(P, oi, n_base, pos, tag, actb, tr); tag_act = tag; break;
            case OP_DOWN:   n_base = consume5<Q41, OP_DOWN, NBL, TRACE>(P, oi, n_base, pos, tag, actb, tr); tag_x = tag; break;
            default:        n_base = consume5<Q41, OP_OUTPUT, NBL, TRACE>(P, oi, n_base, pos, tag, actb, tr); break;
        }
        if (TRACE && tr) tr[3] = clock64();
        if ((tid & 31) == 0) { const unsigned old = atomicAdd(&op_done, 1u); if (old % (unsigned)kConsumerWarps == (unsigned)kConsumerWarps - 1u) hint_post(P.done + oi); }
    }
}

}  // namespace mk5
}  // namespace mg4

// llama_mega.cuh — the persistent decode megakernel: ONE launch per generated token.
//
// Why: a 7B Q4_1 token streams 4.13 GB of weights in ~160 dependent matvecs of 1.6-12 us each; launched one by one the
// HBM pipe drains at every kernel boundary (launch gap + prologue + tail) and never gets above ~25 % of peak.  Weights do
// not depend on activations, so this kernel keeps the weight stream running ACROSS op boundaries:
//   * grid = one CTA per SM (cooperative launch), 16 warps (4 per SM sub-partition, 128 registers each): warp 15 is the
//     PRODUCER, warps 0-14 are CONSUMERS (warps 0-7 also stage activations and run the attention op)
//   * the producer walks the whole op program of the token and copies this CTA's share of every weight matrix
//     HBM -> shared-memory ring with cp.async.bulk (1-D TMA) + mbarrier complete_tx; it runs ahead of the consumers by
//     the ring depth (~200 KB per SM = ~30 MB in flight chip-wide), straight through grid barriers and the attention op
//   * consumers: per op [grid barrier -> stage activations (RMSNorm + Q8 quantise) -> per ring slot: dp4a dot products
//     with warp-shuffle reductions -> fused epilogue], releasing each slot back to the producer through an mbarrier
//   * ops communicate through L2 (x, q, att, act, KV cache) with ld.global.cg loads; a grid barrier (atomic counter)
//     separates dependent ops
// Reduction orders are the canonical ones of llama_kernels.cuh / oracle.cpp, so results are bit-identical to the
// stand-alone kernels and to the CPU oracle.
#pragma once
#include "llama_kernels.cuh"

namespace mg4 {
namespace mk {
using namespace k;

enum OpKind : int { OP_EMBED = 0, OP_QKV = 1, OP_ATTN = 2, OP_WO = 3, OP_GATEUP = 4, OP_DOWN = 5, OP_OUTPUT = 6, OP_FINAL = 7 };

struct MegaOp {
    int kind, layer, rows, cols;
    int row_bytes, sps, n_su, n_warps;   // su = one ROW PAIR, the unit a consumer warp works on; it occupies sps ring slots (1 slot of two
                                         // rows for n_embd-wide matrices, 2 slots of one row for n_ff-wide ones); n_warps = active consumers
    const unsigned char *w;              // row-packed Q4 weights (null for non-matvec ops)
    const float *norm_w;
};

struct MegaParams {
    const MegaOp *ops; int n_ops;
    int n_slots, slot_bytes, act_bytes, xs_bytes;   // shared memory: [ring][act: staged Q8 activations / attention scratch][mbarriers][ops]
    int E, FF, n_head, n_ctx, n_vocab;
    float kq_scale;
    float *x, *q, *att, *act, *logits;
    __half *kcache, *vcache;
    const float2 *rope; const __half *tab_exp, *tab_silu;
    const unsigned char *tok; int tok_type; size_t tok_row_bytes;
    DeviceState *state; unsigned *barrier;
    int l2_ahead;      // producer: ring slots requested into L2 ahead of the fill cursor (0 = off)
    int flags;         // bit 0: request the head's K/V history into L2 (evict_last) while the qkv weights are consumed
                       // bit 1: threads that idle during the attention op request the rest of their gate/up share into L2
    long long *trace;  // optional [2 CTAs][n_ops][8]: clock64 at op start, barrier passed, activations staged, op done; then (warp 0)
                       // cycles spent waiting for ring fills, cycles in the dot products, units processed, unused
};

constexpr int kConsumerWarps = 15, kConsumerThreads = 480;
constexpr int kMegaThreads = kConsumerThreads + 32;  // 15 consumer warps + 1 producer warp = 16 warps
// named barriers: 1 = the 256 threads of warps 0-7 (activation staging, attention); 2 = all 480 consumer threads
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 2, 480;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count)); }
__device__ __forceinline__ void mb_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mb_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory"); }
__device__ __forceinline__ void mb_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MB_DONE;\n\t"
        "bra MB_WAIT;\n\t"
        "MB_DONE:\n\t}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2_keep(const void *p) { asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p)); }
// The attention op of this layer will read K/V rows 0..pos-1 of head h (256 B per key and tensor, written by earlier tokens, long
// since evicted by the weight stream).  Asking L2 for them one op early turns the two HBM round trips on the attention critical path
// into L2 hits; evict_last keeps the weight stream from pushing them out again before they are used.
__device__ __forceinline__ void prefetch_kv_head(const __half *kc, const __half *vc, int pos, int h, int E) {
    for (int i = (int)threadIdx.x; i < 2 * pos; i += 256) {  // (key, 128-byte half of the 256-byte head row)
        const size_t off = (size_t)(i >> 1) * E + h * 128 + (i & 1) * 64;
        prefetch_l2_keep(kc + off); prefetch_l2_keep(vc + off);
    }
}

// first unit (row pair) of CTA `cta`: units are split evenly over the grid (n_su * G < 2^31: host-checked)
__device__ __forceinline__ int unit_begin(int cta, int n_su, int G) { return (int)((unsigned)cta * (unsigned)n_su / (unsigned)G); }

// all consumer threads of all CTAs; `target` = number of arrivals that complete this barrier (monotonic counter)
__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned target) {
    consumer_sync();  // all consumer warps of this CTA have issued their global writes (CTA-scope ordering)
    if (threadIdx.x == 0) {
        // release: a cumulative acq_rel fence publishes the CTA's writes (ordered before it by the bar.sync above) at gpu scope before the
        // arrival.  (__threadfence() is the sequentially-consistent MEMBAR.SC.GPU - stronger, and slower, than this pattern needs.)
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
    }
    consumer_sync();
}

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {  // unsigned bytes x signed bytes
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// sum_j q_j a_j of one 32-weight Q4 block (16 B of nibbles, q in 0..15) with its Q8 block (two 16-byte planes).  The high nibbles
// stay in place (w & 0xF0F0F0F0 = 16 q): their dp4a sum is an exact multiple of 16, so one shift replaces four - and the two
// 4-deep dp4a chains are independent.  Integer arithmetic: same value as the nibble-by-nibble form of k::dot2_q4.
__device__ __forceinline__ int q4_block_idot(const uint4 q, const int4 lo, const int4 hi) {
    int sl = dp4a_us(q.x & 0x0F0F0F0Fu, lo.x, 0), sh = dp4a_us(q.x & 0xF0F0F0F0u, hi.x, 0);
    sl = dp4a_us(q.y & 0x0F0F0F0Fu, lo.y, sl); sh = dp4a_us(q.y & 0xF0F0F0F0u, hi.y, sh);
    sl = dp4a_us(q.z & 0x0F0F0F0Fu, lo.z, sl); sh = dp4a_us(q.z & 0xF0F0F0F0u, hi.z, sh);
    sl = dp4a_us(q.w & 0x0F0F0F0Fu, lo.w, sl); sh = dp4a_us(q.w & 0xF0F0F0F0u, hi.w, sh);
    return sl + (sh >> 4);
}

// Two rows of a shared-memory ring slot against the staged activation vector.  Same per-lane block order and the same
// float arithmetic as k::dot2_q4 (lane l: blocks l, l+32, ... increasing; butterfly reductions), one block per iteration
// (4 independent dp4a chains; shared-memory latency is short and the other warps of the sub-partition fill the gaps, so deeper
// unrolling buys nothing and costs registers).  Q4_0's "-8" is applied as  sum (q-8) a = sum q a - 8 sum a  with the integer
// activation sum the staging pass leaves in the `s` plane.
template <bool Q41>
__device__ __forceinline__ void dot2_q4_slot(const unsigned char *row0, const unsigned char *row1, int nb, int cols, const unsigned char *act, int lane, float &r0, float &r1) {
    const uint4 *qs0 = (const uint4 *)row0, *qs1 = (const uint4 *)row1;
    const unsigned char *sc0 = row0 + (size_t)nb * 16, *sc1 = row1 + (size_t)nb * 16;
    const int4 *alo = (const int4 *)act, *ahi = (const int4 *)(act + cols / 2);
    const float *ad = (const float *)(act + cols), *as = ad + nb;
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

// Activation staging for the megakernel (Q8_0 / Q8_1 targets).  The F32 input vector is read straight from L2 into registers
// (ld.global.cg: other CTAs wrote it earlier in this launch), all loads of a thread in flight at once, and quantised from there
// into the split-plane Q8 layout in shared memory.  A 32-weight quant block is covered by 8 consecutive lanes, so amax / sum need
// 3 shuffle steps; quantisation is order-free (max, integer sums): identical bytes to k::stage_act.
constexpr int kNormItems = 5;   // RMS-normed inputs (warps 0-7): n_embd <= 256 threads x 4 floats x 5 = 5120
constexpr int kPlainItems = 8;  // un-normed inputs (all 15 consumer warps): cols <= 480 x 4 x 8 = 15360
// quantise the float4 at elements i..i+3; `valid` is shared by the 8 lanes of a block (all lanes run the shuffles)
template <int ACT>
__device__ __forceinline__ void quant_item(const float4 a, int i, bool valid, int cols, unsigned char *sm) {
    unsigned char *qs = sm; float *d = (float *)(sm + cols); float *s = d + cols / 32;
    const int j8 = threadIdx.x & 7, b = i >> 5;
    float amax = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4)); amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2)); amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    const float dd = amax / 127.f;
    const float id = amax != 0.0f ? 127.f / amax : 0.0f;
    const int q0 = __float2int_rn(a.x * id), q1 = __float2int_rn(a.y * id), q2 = __float2int_rn(a.z * id), q3 = __float2int_rn(a.w * id);
    int sum = (q0 + q1) + (q2 + q3);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4); sum += __shfl_xor_sync(0xffffffffu, sum, 2); sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    if (valid) {
        *(unsigned *)(qs + (j8 < 4 ? 0 : cols / 2) + b * 16 + (j8 & 3) * 4) = (unsigned)(q0 & 0xff) | ((unsigned)(q1 & 0xff) << 8) | ((unsigned)(q2 & 0xff) << 16) | ((unsigned)(q3 & 0xff) << 24);
        if (j8 == 0) {
            if (ACT == ACT_Q8_0) { d[b] = __half2float(__float2half_rn(dd)); s[b] = (float)sum; }  // integer sum of the block (exact), see dot2_q4_slot
            else { d[b] = dd; s[b] = dd * (float)sum; }
        }
    }
}
// un-normed inputs (wo, down), all 480 consumer threads: thread t owns the float4s 480 k + t
template <int ACT>
__device__ __forceinline__ void stage_plain_mega(const float *__restrict__ x, int cols, unsigned char *sm) {
    const int tid = threadIdx.x;  // tid < 480
    float4 xv[kPlainItems];
#pragma unroll
    for (int it = 0; it < kPlainItems; ++it) {
        const int i = 4 * (tid + kConsumerThreads * it);
        xv[it] = i < cols ? __ldcg((const float4 *)(x + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int it = 0; it < kPlainItems; ++it) {
        const int i = 4 * (tid + kConsumerThreads * it);
        if (4 * kConsumerThreads * it < cols) quant_item<ACT>(xv[it], i, i < cols, cols, sm);  // (CTA-uniform predicate)
    }
}
// RMS-normed inputs (n_embd wide), warps 0-7.  Thread t owns the float4s at elements 1024 it + 4 t: its RMS partials are the
// canonical partials t (even it) and t + 256 (odd it) of oracle.cpp op_rms_norm_mul, same as k::stage_act.  The norm weights
// `nw` were fetched into registers BEFORE the grid barrier (they are static), so only the activations are on the critical path.
__device__ __forceinline__ void load_norm_weights(const float *__restrict__ nw, int cols, float4 (&w)[kNormItems]) {
#pragma unroll
    for (int it = 0; it < kNormItems; ++it) {
        const int i = 1024 * it + 4 * (int)threadIdx.x;
        w[it] = i < cols ? __ldg((const float4 *)(nw + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}
template <int ACT>
__device__ __forceinline__ void stage_norm_mega(const float *__restrict__ x, const float4 (&w)[kNormItems], int cols, unsigned char *sm, double *red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;  // tid < 256
    float4 xv[kNormItems];
#pragma unroll
    for (int it = 0; it < kNormItems; ++it) {
        const int i = 1024 * it + 4 * tid;
        xv[it] = i < cols ? __ldcg((const float4 *)(x + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
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
            quant_item<ACT>(make_float4((a.x * scale) * w4.x, (a.y * scale) * w4.y, (a.z * scale) * w4.z, (a.w * scale) * w4.w), i, i < cols, cols, sm);
        }
    }
}

// out-of-line so that the attention code gets its own register allocation (it runs on n_head CTAs only)
__device__ __noinline__ void attention_mega(const float *q, const __half *kc, const __half *vc, float *out, int pos, int h, int E, int n_ctx, float kq_scale,
                                            const __half *tab_exp, unsigned char *dyn, double *red, float *redf, __half *qh, float *part) {
    attention_head<true>(q, kc, vc, out, pos, h, 0, E, n_ctx, kq_scale, tab_exp, dyn, red, redf, qh, part);
}

// shared memory carve-up (dynamic): [ring][act: staged Q8 activations / attention scratch][full mbarriers][empty mbarriers][ops][fill_count]
struct MegaSmem { unsigned char *ring, *actb; uint64_t *full, *empty; MegaOp *ops; volatile unsigned *fill_count; };
__device__ __forceinline__ MegaSmem carve_smem(const MegaParams &P) {
    extern __shared__ __align__(128) unsigned char smem[];
    MegaSmem m;
    m.ring = smem; m.actb = smem + (size_t)P.n_slots * P.slot_bytes;
    m.full = (uint64_t *)(m.actb + P.act_bytes); m.empty = m.full + P.n_slots;
    m.ops = (MegaOp *)(m.empty + P.n_slots);                       // the op program, copied once from global memory
    m.fill_count = (volatile unsigned *)(m.ops + P.n_ops);         // slots issued so far (read by the L2-prefetch lane)
    return m;
}

// The matvec phase of one op for one consumer warp: for each of its row pairs wait for the ring slot(s), dot, release, epilogue.
// One out-of-line instance per op kind: each gets its own register allocation (the unit loop must not spill - with ~215 KB of
// shared memory the L1 is a few KB, so a spill is an L2 round trip) and only its own epilogue.
//   n_base = ring fill number of this CTA's first unit of the op.  With n_warps * sps <= n_slots a warp can never wait on a slot
//   that is two fills behind (it consumed unit su - n_warps itself), so the mbarrier parity is unambiguous.
template <bool Q41, int KIND, bool TRACE>
__device__ __forceinline__ unsigned consume_units(const MegaParams &P, int oi, unsigned n_base, int pos, long long *tr) {
    const MegaSmem m = carve_smem(P);
    const MegaOp &op = m.ops[oi];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int W = op.n_warps, sps = op.sps;
    const int lo = unit_begin(cta, op.n_su, G), hi = unit_begin(cta + 1, op.n_su, G);
    const unsigned n_next = n_base + (unsigned)(hi - lo) * (unsigned)sps;  // fill number of the next op's first unit
    if (warp >= W) return n_next;
    const int cols = op.cols, nb = cols >> 5, S = P.n_slots, stepn = W * sps;
    const unsigned rb = (unsigned)op.row_bytes, slot_bytes = (unsigned)P.slot_bytes;
    const unsigned n0 = n_base + (unsigned)(warp * sps);
    int s0 = (int)(n0 % (unsigned)S);          // ring slot of this warp's current unit ...
    unsigned ph0 = (n0 / (unsigned)S) & 1u;    // ... and its fill parity (advanced incrementally)
    unsigned long long best = 0ull;
    // gate/up epilogue (lane 0): the SiLU table lookup of a unit is issued after its dot product and consumed after the NEXT
    // unit's, so its L2 latency is off the warp's critical path
    __half pend_h = __ushort_as_half((unsigned short)0); float pend_up = 0.f; int pend_i = -1;
    for (int su = lo + warp; su < hi; su += W) {
        int s1 = s0 + 1; unsigned ph1 = ph0;
        if (s1 == S) { s1 = 0; ph1 ^= 1u; }
        const int r0 = su * 2;
        float2 rs = make_float2(0.f, 0.f);  // residual rows of this pair / RoPE (cos, sin) of this pair: fetched before the wait
        if (KIND == OP_WO || KIND == OP_DOWN) rs = __ldcg((const float2 *)(P.x + r0));
        if (KIND == OP_QKV) { if (r0 < 2 * P.E) rs = __ldg(&P.rope[(size_t)pos * 64 + ((r0 % P.E) % 128) / 2]); }
        long long tw0 = 0, tw1 = 0, tw2 = 0;
        if (TRACE && tr) tw0 = clock64();
        mb_wait(&m.full[s0], ph0);
        if (sps == 2) mb_wait(&m.full[s1], ph1);
        if (TRACE && tr) tw1 = clock64();
        const unsigned char *row0 = m.ring + (size_t)s0 * slot_bytes;
        const unsigned char *row1 = sps == 2 ? m.ring + (size_t)s1 * slot_bytes : row0 + rb;
        float v0, v1;
        dot2_q4_slot<Q41>(row0, row1, nb, cols, m.actb, lane, v0, v1);
        if (TRACE && tr) { tw2 = clock64(); tr[4] += tw1 - tw0; tr[5] += tw2 - tw1; tr[6] += 1; }
        if (lane == 0) {
            mb_arrive(&m.empty[s0]);
            if (sps == 2) mb_arrive(&m.empty[s1]);
            if (KIND == OP_QKV) {
                const int E = P.E, partn = r0 / E, rr = r0 % E;
                const size_t kvo = ((size_t)op.layer * P.n_ctx + pos) * E + rr;
                if (partn == 2) { *(__half2 *)(P.vcache + kvo) = __floats2half2_rn(v0, v1); }
                else {
                    const float2 cs = rs;
                    const float o0 = v0 * cs.x - v1 * cs.y, o1 = v0 * cs.y + v1 * cs.x;
                    if (partn == 0) *(float2 *)(P.q + rr) = make_float2(o0, o1);
                    else *(__half2 *)(P.kcache + kvo) = __floats2half2_rn(o0, o1);
                }
            } else if (KIND == OP_WO || KIND == OP_DOWN) {
                *(float2 *)(P.x + r0) = make_float2(v0 + rs.x, v1 + rs.y);
            } else if (KIND == OP_GATEUP) {
                if (pend_i >= 0) P.act[pend_i] = __half2float(pend_h) * pend_up;
                pend_h = P.tab_silu[__half_as_ushort(__float2half_rn(v0))]; pend_up = v1; pend_i = r0 >> 1;
            } else {  // OP_OUTPUT
                P.logits[r0] = v0;
                const unsigned long long k0 = argmax_key(v0, r0);
                best = best > k0 ? best : k0;
                if (r0 + 1 < P.n_vocab) { P.logits[r0 + 1] = v1; const unsigned long long k1 = argmax_key(v1, r0 + 1); best = best > k1 ? best : k1; }
            }
        }
        s0 += stepn; while (s0 >= S) { s0 -= S; ph0 ^= 1u; }
        if (TRACE && tr) tr[7] += clock64() - tw2;
    }
    if (KIND == OP_GATEUP) { if (lane == 0 && pend_i >= 0) P.act[pend_i] = __half2float(pend_h) * pend_up; }
    if (KIND == OP_OUTPUT) { if (lane == 0 && best) atomicMax(&P.state->argmax_key, best); }
    return n_next;
}

// The front half of a matvec op for the consumer warps: fetch the (static) RMSNorm weights, pass the grid barrier that makes the
// previous op's output visible, stage the input vector as Q8 blocks in shared memory.
template <int ACT, bool TRACE>
__device__ __forceinline__ void stage_op(const MegaParams &P, int oi, unsigned bar_target, long long *tr) {
    __shared__ double red[34];
    const MegaSmem m = carve_smem(P);
    const MegaOp &op = m.ops[oi];
    const int tid = threadIdx.x, cols = op.cols, kind = op.kind;
    const float *nw = op.norm_w;
    // the (static) norm weights are only ASKED into L2 before the barrier and loaded together with the activations after it: holding
    // them in registers across the barrier cost 20 registers that ptxas spilled to local memory (an L2 round trip on this path)
    if (nw && tid * 32 < cols) prefetch_l2(nw + tid * 32);
    grid_barrier(P.barrier, bar_target);
    if (TRACE && tr) tr[1] = clock64();
    const float *src = kind == OP_WO ? P.att : kind == OP_DOWN ? P.act : P.x;
    if (nw) { if (tid < 256) { float4 nwr[kNormItems]; load_norm_weights(nw, cols, nwr); stage_norm_mega<ACT>(src, nwr, cols, m.actb, red); } }
    else stage_plain_mega<ACT>(src, cols, m.actb);
    consumer_sync();
    if (TRACE && tr) tr[2] = clock64();
}

// The producer warp: lane 0 fills the shared-memory ring with cp.async.bulk (blocks when the ring is full); lane 1 walks the same
// slot sequence up to l2_ahead slots further and only asks L2 to fetch (cp.async.bulk.prefetch.L2), so HBM keeps streaming while
// the consumers sit in a grid barrier / staging / the attention op.  Per-op fields are held in registers.
__device__ __noinline__ void producer_loop(const MegaParams &P) {
    const MegaSmem m = carve_smem(P);
    unsigned char *const ring = m.ring;
    uint64_t *const full = m.full, *const empty = m.empty;
    const MegaOp *const ops = m.ops;
    volatile unsigned *const fill_count = m.fill_count;
    const int lane = threadIdx.x & 31, G = (int)gridDim.x, cta = (int)blockIdx.x;
    if (lane == 0) {
        unsigned n = 0, s = 0, ph = 0;  // fill number, its ring slot and phase parity (running counters: no div/mod per slot)
        for (int oi = 0; oi < P.n_ops; ++oi) {
            const unsigned char *w = ops[oi].w;
            if (!w) continue;
            const int n_su = ops[oi].n_su, sps = ops[oi].sps;
            const unsigned rb = (unsigned)ops[oi].row_bytes, bytes = sps == 1 ? 2u * rb : rb;
            const int lo = unit_begin(cta, n_su, G), hi = unit_begin(cta + 1, n_su, G);
            const unsigned char *src = w + (size_t)lo * 2 * rb;
            for (int c = (hi - lo) * sps; c > 0; --c, src += bytes) {
                mb_wait(&empty[s], ph ^ 1u);
                mb_expect_tx(&full[s], bytes);
                bulk_g2s(ring + (size_t)s * P.slot_bytes, src, bytes, &full[s]);
                *fill_count = ++n;
                if (++s == (unsigned)P.n_slots) { s = 0; ph ^= 1u; }
            }
        }
    } else if (lane == 1 && P.l2_ahead > 0) {
        unsigned n = 0;
        for (int oi = 0; oi < P.n_ops; ++oi) {
            const unsigned char *w = ops[oi].w;
            if (!w) continue;
            const int n_su = ops[oi].n_su, sps = ops[oi].sps;
            const unsigned rb = (unsigned)ops[oi].row_bytes;
            const int lo = unit_begin(cta, n_su, G), hi = unit_begin(cta + 1, n_su, G);
            const unsigned char *src = w + (size_t)lo * 2 * rb;
            for (int su = lo; su < hi; ++su, n += (unsigned)sps, src += 2 * rb) {
                unsigned fc;
                while ((int)(n - (fc = *fill_count)) >= P.l2_ahead) __nanosleep(100);
                // only ever ask for data the fill cursor has not reached yet: a request BEHIND the cursor re-reads from HBM what the ring
                // already holds or has consumed (ncu, r1_v3: 8.1 GB of DRAM reads per token for 4.13 GB of weights, L2 hit rate 5 %)
                if ((int)(n - fc) > 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(2u * rb) : "memory");
            }
        }
    }
}

// token embedding row -> x (dequantised on the fly), spread over the grid
__device__ __noinline__ void embed_op(const MegaParams &P) {
    const int token = __ldcg(&P.state->tokens[0]);
    const unsigned char *row = P.tok + (size_t)token * P.tok_row_bytes;
    for (int i = (int)blockIdx.x * kConsumerThreads + (int)threadIdx.x; i < P.E; i += (int)gridDim.x * kConsumerThreads) P.x[i] = dequant_elem(P.tok_type, row, i);
}

template <int WT, bool TRACE>
__global__ void __launch_bounds__(kMegaThreads, 1) decode_megakernel(const __grid_constant__ MegaParams P) {
    __shared__ double red[34];
    __shared__ float redf[34];
    __shared__ __align__(16) __half qh[128];
    __shared__ float part[16 * 128];
    constexpr int ACT = act_of(WT);
    constexpr bool Q41 = WT == GG_Q4_1;
    const MegaSmem m = carve_smem(P);
    unsigned char *const ring = m.ring, *const actb = m.actb;
    uint64_t *const full = m.full, *const empty = m.empty;
    MegaOp *const ops = m.ops;
    volatile unsigned *const fill_count = m.fill_count;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = (int)gridDim.x, cta = (int)blockIdx.x;

    if (tid == 0) {
        for (int s = 0; s < P.n_slots; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); }
        *fill_count = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < P.n_ops * (int)(sizeof(MegaOp) / 16); i += kMegaThreads) ((uint4 *)ops)[i] = ((const uint4 *)P.ops)[i];
    __syncthreads();  // the only CTA-wide barrier; afterwards consumers use named barriers 2 (512 thr) and 1 (256 thr, attention)

    if (warp == kConsumerWarps) { producer_loop(P); return; }

    // ------------------------------------ consumers ------------------------------------
    // The op loop keeps almost nothing live: the phases are out-of-line functions (stage_op, consume_units, attention_mega) that
    // ptxas gives the registers above the caller's live set - the smaller this loop's state, the more the hot loops get.
    unsigned n_base = 0, bar_target = 0;
    const int pos = __ldcg(&P.state->n_past);  // position of the token being decoded (state only changes in OP_FINAL)
    for (int oi = 0; oi < P.n_ops; ++oi) {
        const int kind = ops[oi].kind;
        long long *tr = nullptr;
        if (TRACE) { if (P.trace && tid == 0 && (cta == 0 || cta == G - 1)) tr = P.trace + ((size_t)(cta == 0 ? 0 : 1) * P.n_ops + oi) * 8; }
        if (TRACE && tr) { tr[0] = clock64(); tr[2] = 0; tr[4] = 0; tr[5] = 0; tr[6] = 0; tr[7] = 0; }
        if (oi > 0) bar_target += (unsigned)G;
        if (kind == OP_EMBED || kind == OP_ATTN || kind == OP_FINAL) {
            if (oi > 0) grid_barrier(P.barrier, bar_target);
            if (TRACE && tr) tr[1] = clock64();
            if (kind == OP_EMBED) embed_op(P);
            else if (kind == OP_ATTN) {
                if (cta < P.n_head && tid < 256) {  // one head per CTA, 256 threads (named barrier 1)
                    const size_t lo = (size_t)ops[oi].layer * P.n_ctx * P.E;
                    attention_mega(P.q, P.kcache + lo, P.vcache + lo, P.att, pos, cta, P.E, P.n_ctx, P.kq_scale, P.tab_exp, actb, red, redf, qh, part);
                } else if ((P.flags & 2) && oi + 2 < P.n_ops && ops[oi + 2].w) {
                    // Everybody else idles until the heads are done, the ring is full and the producer is blocked, so HBM idles too.  These
                    // threads ask L2 for the part of this CTA's gate/up share that the ring has not requested yet; the producer's later
                    // fills then come out of L2.  (Requests strictly ahead of the fill cursor: nothing is read from DRAM twice.)
                    const MegaOp &wo = ops[oi + 1], &gu = ops[oi + 2];
                    const unsigned cnt_wo = (unsigned)(unit_begin(cta + 1, wo.n_su, G) - unit_begin(cta, wo.n_su, G)) * (unsigned)wo.sps;
                    const int lo = unit_begin(cta, gu.n_su, G), hi = unit_begin(cta + 1, gu.n_su, G);
                    const int issued = (int)(*fill_count - (n_base + cnt_wo)) / gu.sps;  // gate/up units the producer has already requested
                    const int first = lo + (issued > 0 ? issued : 0);
                    if (first < hi) {
                        const unsigned char *base = gu.w + (size_t)first * 2 * gu.row_bytes;
                        const size_t bytes = (size_t)(hi - first) * 2 * gu.row_bytes;
                        const int t = cta < P.n_head ? tid - 256 : tid, nt = cta < P.n_head ? kConsumerThreads - 256 : kConsumerThreads;
                        for (size_t off = (size_t)t * 128; off < bytes; off += (size_t)nt * 128) prefetch_l2(base + off);
                    }
                }
            } else if (cta == 0 && tid == 0) {
                DeviceState *st = P.state;
                const unsigned long long key = __ldcg((const unsigned long long *)&st->argmax_key);
                const int id = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
                st->argmax_id = id; st->tokens[0] = id; st->argmax_key = 0ull;
                st->n_past += 1; st->n_tok = 1;
            }
            continue;
        }
        // ---- matvec ops: [norm weights -> grid barrier -> stage activations] then [ring slots -> dot -> epilogue] ----
        stage_op<ACT, TRACE>(P, oi, bar_target, tr);
        if (kind == OP_QKV && (P.flags & 1) && cta < P.n_head && tid < 256) {
            const size_t lo = (size_t)ops[oi].layer * P.n_ctx * P.E;
            prefetch_kv_head(P.kcache + lo, P.vcache + lo, pos, cta, P.E);
        }
        switch (kind) {
            case OP_QKV:    n_base = consume_units<Q41, OP_QKV, TRACE>(P, oi, n_base, pos, tr); break;
            case OP_WO:     n_base = consume_units<Q41, OP_WO, TRACE>(P, oi, n_base, pos, tr); break;
            case OP_GATEUP: n_base = consume_units<Q41, OP_GATEUP, TRACE>(P, oi, n_base, pos, tr); break;
            case OP_DOWN:   n_base = consume_units<Q41, OP_DOWN, TRACE>(P, oi, n_base, pos, tr); break;
            default:        n_base = consume_units<Q41, OP_OUTPUT, TRACE>(P, oi, n_base, pos, tr); break;
        }
        if (TRACE && tr) tr[3] = clock64();  // (thread 0 = warp 0 only; other warps may still be consuming)
    }
}

}  // namespace mk
}  // namespace mg4

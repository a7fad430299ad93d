// llama_mega_ll.cuh — EXPERIMENTAL variant of the decode megakernel (enabled with MINIGPT4_B200_MEGA_LL=1; NOT the default and not yet
// measured: written at the end of round 1 after the GPU budget was spent, see DESIGN.md §7 "Round-2 plan (a)").
//
// Why: in the default kernel (llama_mega.cuh) 19.5 of 45 us per layer are "grid barrier + reload": the last producer's fence + atomic,
// the pollers' round trip, a CTA barrier, and only THEN the loads of the activation vector (r1_v4_mega_trace.txt).  Here the activation
// vectors that ops exchange (x, q, att, act, and the current position's K / V) carry their own readiness: every element is an
// 8-byte {value, tag} pair written and read as ONE 64-bit access (NCCL's "LL" protocol), tag = (launch sequence << 10) + op index + 1,
// so a consumer's load IS its barrier - it spins on the elements it needs until their tags match the op it is staging.  No grid
// barrier is left between the ops of a layer; one remains in front of OP_FINAL (the arg-max of all logits).
//
// Why no write-after-read hazard appears without the barriers: every consumer of a vector gathers ALL of its elements, and every
// producer of the next version of a vector transitively depends on such a full gather (x' needs all of att, att needs all of q/k/v,
// q/k/v need all of x, ...), so all reads of version n have completed on every CTA before any CTA can produce an element of version n+1.
// Tags are unique per (launch, op), hence a stale element can only ever look "not ready", never "ready".
//
// Arithmetic, reduction orders and the ring / producer warp are those of llama_mega.cuh (its helpers are reused unchanged).
#pragma once
#include "llama_mega.cuh"

namespace mg4 {
namespace mk {

struct LLf { float v; unsigned tag; };  // 8 bytes, 8-byte aligned: one 64-bit access

struct MegaLL {             // extra buffers of the LL variant (device pointers), appended to the launch parameters
    LLf *x, *q, *att, *act; // [E], [E], [E], [FF]
    LLf *kcur, *vcur;       // [E/2] each: the current position's K / V as {half2 bits, tag} (the F16 cache rows have no room for a tag)
    unsigned *seq;          // launches so far (advanced by OP_FINAL); part of every tag
};
struct MegaParamsLL { MegaParams p; MegaLL ll; };

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
// Every spin is bounded: a protocol bug must end in a trapped launch (an error the host reports), never in a hung GPU.
constexpr unsigned kLLSpinLimit = 1u << 22;  // ~ seconds of polling; a healthy wait is microseconds
__device__ __forceinline__ void ll_spin_guard(unsigned &spins) { if (++spins > kLLSpinLimit) asm volatile("trap;"); }
// spin until the element carries `tag`; returns the payload bits
__device__ __forceinline__ unsigned ll_wait(const LLf *p, unsigned tag) {
    unsigned long long w; unsigned spins = 0;
    do { w = ll_load(p); ll_spin_guard(spins); } while ((unsigned)(w >> 32) != tag);
    return (unsigned)w;
}
// two elements with one 16-byte access (each 8-byte half is written by one 64-bit store, so it is seen whole - the access pattern of NCCL's
// LL128 reads); halves the L2->SM requests of a polling round compared with four 8-byte loads per float4
__device__ __forceinline__ void ll_load2(const LLf *p, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
// four consecutive elements (one float4 of payload, 32-byte aligned); both loads are in flight before the first tag is looked at
__device__ __forceinline__ bool ll_try4(const LLf *p, unsigned tag, float4 &out) {
    unsigned long long a, b, c, d;
    ll_load2(p, a, b); ll_load2(p + 2, c, d);
    out = make_float4(__uint_as_float((unsigned)a), __uint_as_float((unsigned)b), __uint_as_float((unsigned)c), __uint_as_float((unsigned)d));
    return (unsigned)(a >> 32) == tag && (unsigned)(b >> 32) == tag && (unsigned)(c >> 32) == tag && (unsigned)(d >> 32) == tag;
}

// un-normed inputs (wo <- att, down <- act), all 480 consumer threads; same element ownership as stage_plain_mega
template <int ACT>
__device__ __forceinline__ void stage_plain_ll(const LLf *__restrict__ x, unsigned tag, int cols, unsigned char *sm, long long *tr) {
    const int tid = threadIdx.x;
    float4 xv[kPlainItems];
    unsigned need = 0;
#pragma unroll
    for (int it = 0; it < kPlainItems; ++it) { xv[it] = make_float4(0.f, 0.f, 0.f, 0.f); if (4 * (tid + kConsumerThreads * it) < cols) need |= 1u << it; }
    unsigned spins = 0;
    while (need) {
        ll_spin_guard(spins);
#pragma unroll
        for (int it = 0; it < kPlainItems; ++it)
            if (need & (1u << it)) { if (ll_try4(x + 4 * (tid + kConsumerThreads * it), tag, xv[it])) need &= ~(1u << it); }
    }
    if (tr) tr[1] = clock64();  // this thread's share of the input has arrived
#pragma unroll
    for (int it = 0; it < kPlainItems; ++it) {
        const int i = 4 * (tid + kConsumerThreads * it);
        if (4 * kConsumerThreads * it < cols) quant_item<ACT>(xv[it], i, i < cols, cols, sm);  // (CTA-uniform predicate)
    }
}
// RMS-normed inputs (qkv, gate/up, output <- x), warps 0-7; same element ownership and reduction order as stage_norm_mega
template <int ACT>
__device__ __forceinline__ void stage_norm_ll(const LLf *__restrict__ x, unsigned tag, const float *__restrict__ nw, int cols, unsigned char *sm, double *red, long long *tr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;  // tid < 256
    float4 xv[kNormItems];
    unsigned need = 0;
#pragma unroll
    for (int it = 0; it < kNormItems; ++it) { xv[it] = make_float4(0.f, 0.f, 0.f, 0.f); if (1024 * it + 4 * tid < cols) need |= 1u << it; }
    unsigned spins = 0;
    while (need) {
        ll_spin_guard(spins);
#pragma unroll
        for (int it = 0; it < kNormItems; ++it)
            if (need & (1u << it)) { if (ll_try4(x + 1024 * it + 4 * tid, tag, xv[it])) need &= ~(1u << it); }
    }
    if (tr) tr[1] = clock64();  // this thread's share of the input has arrived
    float4 w[kNormItems];
    load_norm_weights(nw, cols, w);
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

// front half of a matvec op: wait (on the data itself) for the input vector, stage it as Q8 blocks in shared memory
template <int ACT>
__device__ __forceinline__ void stage_op_ll(const MegaParamsLL &PL, int oi, unsigned tag_in, long long *tr) {
    __shared__ double red[34];
    const MegaParams &P = PL.p;
    const MegaSmem m = carve_smem(P);
    const MegaOp &op = m.ops[oi];
    const int tid = threadIdx.x, cols = op.cols, kind = op.kind;
    const float *nw = op.norm_w;
    if (nw && tid * 32 < cols) prefetch_l2(nw + tid * 32);
    consumer_sync();  // every consumer warp of this CTA has finished the previous op: the staging area may be overwritten
    const LLf *src = kind == OP_WO ? PL.ll.att : kind == OP_DOWN ? PL.ll.act : PL.ll.x;
    if (nw) { if (tid < 256) stage_norm_ll<ACT>(src, tag_in, nw, cols, m.actb, red, tr); }
    else stage_plain_ll<ACT>(src, tag_in, cols, m.actb, tr);
    consumer_sync();
    if (tr) tr[2] = clock64();
}

// matvec phase of one op: as consume_units, with tagged outputs
template <bool Q41, int KIND>
__device__ __forceinline__ unsigned consume_units_ll(const MegaParamsLL &PL, int oi, unsigned n_base, int pos, unsigned tag) {
    const MegaParams &P = PL.p;
    const MegaSmem m = carve_smem(P);
    const MegaOp &op = m.ops[oi];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, G = (int)gridDim.x, cta = (int)blockIdx.x;
    const int W = op.n_warps, sps = op.sps;
    const int lo = unit_begin(cta, op.n_su, G), hi = unit_begin(cta + 1, op.n_su, G);
    const unsigned n_next = n_base + (unsigned)(hi - lo) * (unsigned)sps;
    if (warp >= W) return n_next;
    const int cols = op.cols, nb = cols >> 5, S = P.n_slots, stepn = W * sps;
    const unsigned rb = (unsigned)op.row_bytes, slot_bytes = (unsigned)P.slot_bytes;
    const unsigned n0 = n_base + (unsigned)(warp * sps);
    int s0 = (int)(n0 % (unsigned)S);
    unsigned ph0 = (n0 / (unsigned)S) & 1u;
    unsigned long long best = 0ull;
    __half pend_h = __ushort_as_half((unsigned short)0); float pend_up = 0.f; int pend_i = -1;
    for (int su = lo + warp; su < hi; su += W) {
        int s1 = s0 + 1; unsigned ph1 = ph0;
        if (s1 == S) { s1 = 0; ph1 ^= 1u; }
        const int r0 = su * 2;
        float2 rs = make_float2(0.f, 0.f);
        if (KIND == OP_WO || KIND == OP_DOWN) {  // residual rows: final since this CTA staged the whole of x for the previous normed op
            const uint4 t = __ldcg((const uint4 *)(PL.ll.x + r0));
            rs = make_float2(__uint_as_float(t.x), __uint_as_float(t.z));
        }
        if (KIND == OP_QKV) { if (r0 < 2 * P.E) rs = __ldg(&P.rope[(size_t)pos * 64 + ((r0 % P.E) % 128) / 2]); }
        mb_wait(&m.full[s0], ph0);
        if (sps == 2) mb_wait(&m.full[s1], ph1);
        const unsigned char *row0 = m.ring + (size_t)s0 * slot_bytes;
        const unsigned char *row1 = sps == 2 ? m.ring + (size_t)s1 * slot_bytes : row0 + rb;
        float v0, v1;
        dot2_q4_slot<Q41>(row0, row1, nb, cols, m.actb, lane, v0, v1);
        if (lane == 0) {
            mb_arrive(&m.empty[s0]);
            if (sps == 2) mb_arrive(&m.empty[s1]);
            if (KIND == OP_QKV) {
                const int E = P.E, partn = r0 / E, rr = r0 % E;
                const size_t kvo = ((size_t)op.layer * P.n_ctx + pos) * E + rr;
                if (partn == 2) {
                    const __half2 h2 = __floats2half2_rn(v0, v1);
                    *(__half2 *)(P.vcache + kvo) = h2;                                   // for later tokens (stream order)
                    ll_store_bits(PL.ll.vcur + (rr >> 1), *(const unsigned *)&h2, tag);  // for this token's attention
                } else {
                    const float2 cs = rs;
                    const float o0 = v0 * cs.x - v1 * cs.y, o1 = v0 * cs.y + v1 * cs.x;
                    if (partn == 0) { ll_store(PL.ll.q + rr, o0, tag); ll_store(PL.ll.q + rr + 1, o1, tag); }
                    else {
                        const __half2 h2 = __floats2half2_rn(o0, o1);
                        *(__half2 *)(P.kcache + kvo) = h2;
                        ll_store_bits(PL.ll.kcur + (rr >> 1), *(const unsigned *)&h2, tag);
                    }
                }
            } else if (KIND == OP_WO || KIND == OP_DOWN) {
                ll_store(PL.ll.x + r0, v0 + rs.x, tag); ll_store(PL.ll.x + r0 + 1, v1 + rs.y, tag);
            } else if (KIND == OP_GATEUP) {
                if (pend_i >= 0) ll_store(PL.ll.act + pend_i, __half2float(pend_h) * pend_up, tag);
                pend_h = P.tab_silu[__half_as_ushort(__float2half_rn(v0))]; pend_up = v1; pend_i = r0 >> 1;
            } else {  // OP_OUTPUT
                P.logits[r0] = v0;
                const unsigned long long k0 = argmax_key(v0, r0);
                best = best > k0 ? best : k0;
                if (r0 + 1 < P.n_vocab) { P.logits[r0 + 1] = v1; const unsigned long long k1 = argmax_key(v1, r0 + 1); best = best > k1 ? best : k1; }
            }
        }
        s0 += stepn; while (s0 >= S) { s0 -= S; ph0 ^= 1u; }
    }
    if (KIND == OP_GATEUP) { if (lane == 0 && pend_i >= 0) ll_store(PL.ll.act + pend_i, __half2float(pend_h) * pend_up, tag); }
    if (KIND == OP_OUTPUT) { if (lane == 0 && best) atomicMax(&P.state->argmax_key, best); }
    return n_next;
}

// TRACE: per-op clock stamps of thread 0 of CTA 0 and CTA G-1 in the layout tools/mega_trace.py reads ({op start, input arrived, staged, done};
// the "barrier" column then means "waiting for the input data"); the per-unit fields stay zero
template <int WT, bool TRACE>
__global__ void __launch_bounds__(kMegaThreads, 1) decode_megakernel_ll(const __grid_constant__ MegaParamsLL PL) {
    __shared__ double red[34];
    __shared__ float redf[34];
    __shared__ __align__(16) __half qh[128];
    __shared__ float part[16 * 128];
    constexpr int ACT = act_of(WT);
    constexpr bool Q41 = WT == GG_Q4_1;
    const MegaParams &P = PL.p;
    const MegaSmem m = carve_smem(P);
    uint64_t *const full = m.full, *const empty = m.empty;
    MegaOp *const ops = m.ops;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int G = (int)gridDim.x, cta = (int)blockIdx.x;

    if (tid == 0) {
        for (int s = 0; s < P.n_slots; ++s) { mb_init(&full[s], 1); mb_init(&empty[s], 1); }
        *m.fill_count = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < P.n_ops * (int)(sizeof(MegaOp) / 16); i += kMegaThreads) ((uint4 *)ops)[i] = ((const uint4 *)P.ops)[i];
    __syncthreads();

    if (warp == kConsumerWarps) { producer_loop(P); return; }

    unsigned n_base = 0;
    const int pos = __ldcg(&P.state->n_past);
    const unsigned tag0 = (__ldcg(PL.ll.seq) << 10) + 1u;  // tag of op oi in this launch = tag0 + oi
    unsigned tag_x = 0, tag_att = 0, tag_act = 0, tag_qkv = 0;  // tag of the op that last produced each vector
    for (int oi = 0; oi < P.n_ops; ++oi) {
        const int kind = ops[oi].kind;
        const unsigned tag = tag0 + (unsigned)oi;
        long long *tr = nullptr;
        if (TRACE) { if (P.trace && tid == 0 && (cta == 0 || cta == G - 1)) tr = P.trace + ((size_t)(cta == 0 ? 0 : 1) * P.n_ops + oi) * 8; }
        if (TRACE && tr) { tr[0] = clock64(); tr[1] = tr[0]; tr[2] = 0; tr[3] = 0; tr[4] = 0; tr[5] = 0; tr[6] = 0; tr[7] = 0; }
        if (kind == OP_EMBED) {
            const int token = __ldcg(&P.state->tokens[0]);
            const unsigned char *row = P.tok + (size_t)token * P.tok_row_bytes;
            for (int i = cta * kConsumerThreads + tid; i < P.E; i += G * kConsumerThreads) ll_store(PL.ll.x + i, dequant_elem(P.tok_type, row, i), tag);
            tag_x = tag;
            continue;
        }
        if (kind == OP_ATTN) {
            if (cta < P.n_head) {
                consumer_sync();  // all 15 warps are done with the qkv rows: the staging area becomes attention scratch
                if (tid < 256) {
                    // gather this head's q and the current position's K / V (tagged by the qkv op) into the plain buffers attention_head reads
                    const int h = cta, E = P.E;
                    const size_t lo = (size_t)ops[oi].layer * P.n_ctx * E;
                    if (tid < 128) P.q[h * 128 + tid] = __uint_as_float(ll_wait(PL.ll.q + h * 128 + tid, tag_qkv));
                    else if (tid < 192) { const int j = tid - 128; *(unsigned *)(P.kcache + lo + (size_t)pos * E + h * 128 + 2 * j) = ll_wait(PL.ll.kcur + h * 64 + j, tag_qkv); }
                    else { const int j = tid - 192; *(unsigned *)(P.vcache + lo + (size_t)pos * E + h * 128 + 2 * j) = ll_wait(PL.ll.vcur + h * 64 + j, tag_qkv); }
                    __threadfence_block();
                    cta_sync<true>();
                    if (TRACE && tr) tr[1] = clock64();
                    attention_mega(P.q, P.kcache + lo, P.vcache + lo, P.att, pos, h, E, P.n_ctx, P.kq_scale, P.tab_exp, m.actb, red, redf, qh, part);
                    if (tid < 128) ll_store(PL.ll.att + h * 128 + tid, __ldcg(&P.att[h * 128 + tid]), tag);  // (each thread re-reads its own store)
                }
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
                *PL.ll.seq = (tag0 >> 10) + 1u;
            }
            continue;
        }
        // ---- matvec ops ----
        const unsigned tag_in = kind == OP_WO ? tag_att : kind == OP_DOWN ? tag_act : tag_x;
        stage_op_ll<ACT>(PL, oi, tag_in, TRACE ? tr : nullptr);
        if (kind == OP_QKV && (P.flags & 1) && cta < P.n_head && tid < 256) {
            const size_t lo = (size_t)ops[oi].layer * P.n_ctx * P.E;
            prefetch_kv_head(P.kcache + lo, P.vcache + lo, pos, cta, P.E);
        }
        switch (kind) {
            case OP_QKV:    n_base = consume_units_ll<Q41, OP_QKV>(PL, oi, n_base, pos, tag); tag_qkv = tag; break;
            case OP_WO:     n_base = consume_units_ll<Q41, OP_WO>(PL, oi, n_base, pos, tag); tag_x = tag; break;
            case OP_GATEUP: n_base = consume_units_ll<Q41, OP_GATEUP>(PL, oi, n_base, pos, tag); tag_act = tag; break;
            case OP_DOWN:   n_base = consume_units_ll<Q41, OP_DOWN>(PL, oi, n_base, pos, tag); tag_x = tag; break;
            default:        n_base = consume_units_ll<Q41, OP_OUTPUT>(PL, oi, n_base, pos, tag); break;
        }
        if (TRACE && tr) tr[3] = clock64();
    }
}

}  // namespace mk
}  // namespace mg4

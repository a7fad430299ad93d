// llama_prefill.cuh — tensor-core prefill for Q4_0 / Q4_1 LLaMA matrices: N >= 2 rows of llama_eval / llama_eval_embd in ONE pass over
// the weights (reference call sites minigpt4.cpp:2365-2382 add_tokens in n_batch chunks, :2399-2415 add_embedding = 32 rows at once).
//
// Semantics are ggml's quantised mul_mat, exactly as the decode matvec (llama_kernels.cuh): activations quantised to Q8_0 / Q8_1 per 32,
// INTEGER block dot products, F32 scaling.  Here the integer dots run on the 5th-generation tensor cores:
//   * one tcgen05.mma kind::i8 (M = 128 weight rows, N = 32 tokens, K = 32 = ONE quant block, accumulate = 0) per quant block puts the
//     128 x 32 exact int32 block dots into TMEM; the per-block scales cannot be folded into the MMA, so
//   * sixteen epilogue warps (32 rows x 8 tokens each) read each block's dots back (tcgen05.ld), apply  d_w * d_a * isum  and  m_w * s_a  on the CUDA cores and
//     accumulate in the CANONICAL FLOAT ORDER of oracle.cpp / k::dot2_q4: "lane" class l = b mod 32 accumulates its blocks l, l+32, ...
//     in increasing order; the 32 class sums are then combined by the xor-butterfly tree (16, 8, 4, 2, 1).  The K loop therefore walks
//     the classes in bit-reversed order (0, 16, 8, 24, ...), which turns the butterfly into a post-order tree walk with a 5-deep stack
//     per output (level 0 in registers, levels 1-4 in shared memory).  Result: bit-identical to the matvec path, for any N.
//   * operands come from a load-time "prefill operand cache": the nibbles expanded to int8 (Q4_0: q - 8, so the "-8 * sum(a)" term is
//     inside the integer dot) in CLASS-MAJOR order — row r = [class l][slot i][32 bytes] — so that one 128-byte TMA box row holds four
//     consecutive blocks of one class (SWIZZLE_128B, the layout the tensor core reads); the staged activations use the same order.
//     (Trade: +1 byte per weight of HBM for the cache on a 180 GB device; the decode path keeps streaming the 4-bit original.)
// The kernel is epilogue-bound by construction (5 CUDA-core instructions per (row, token, block) against 32 int8 MACs on the tensor core):
// exact Q8 semantics cost that; what the tensor core removes is the 8 dp4a + nibble unpacking per (row, token, block) of the matvec path.
#pragma once
#include "llama_kernels.cuh"
#include <cuda.h>

namespace mg4 {
namespace pf {
using namespace k;

constexpr int kTok = 32;          // tokens per CTA (UMMA N)
constexpr int kRows = 128;        // weight rows per CTA (UMMA M)
constexpr int kStages = 4;        // shared-memory pipeline depth (tiles of 4 blocks)
constexpr int kStageBytes = 16384 + 4096 + 1024;  // A 128 x 128 B | B 32 x 128 B | activation scales 32 x 4 x {d, s}
constexpr int kGroups = 4;        // TMEM tile groups (4 x 32-column block regions each) -> 512 columns
constexpr int kEpiWarps = 16;       // 4 per TMEM lane quarter: a warp owns 32 rows x kTpw tokens (twice the warps of the first cut: the epilogue is a latency chain)
constexpr int kTpw = kTok * 4 / kEpiWarps;   // tokens per epilogue warp (8)
constexpr int kThreads = 64 + 32 * kEpiWarps;     // warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-17 = epilogue
static_assert(kTpw == 8, "tmem_ld8i");
constexpr int kStackBytes = 4 * kTok * kRows * 8; // stack levels 1-4: [level][token][row] {d-tree, m-tree}

struct PrefillArgs {
    int rows;            // valid weight rows (the operand cache is padded to a multiple of 128 rows)
    int n_tok;           // valid tokens (grid.y = ceil(n_tok / 32))
    int nb, S;           // quant blocks per row; slots per class in the class-major layout (multiple of 4)
    const __half2 *wsc;  // [rows_pad][32 * S] {d, m} of the block in (class, slot) order
    int epi;             // k::Epi (EPI_PLAIN, EPI_QKV, EPI_RESID, EPI_SWIGLU)
    float *out; int out_stride; const float *resid;
    float *q_out; __half *kcache; __half *vcache; const float2 *rope; int e_local; int half_dim;
    const DeviceState *state;
    const __half *tab_silu;
    // K split (grid.z > 1): slice z covers the classes of the leaves jr in [z * jr_per_z, (z + 1) * jr_per_z) of the butterfly tree - a complete
    // subtree - and writes its root {d-tree, m-tree} to partial[(z * part_tok + t) * part_rows + r]; prefill_combine folds the roots and runs the epilogue
    int jr_per_z; float2 *partial; int part_tok, part_rows;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "PF_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PF_DONE;\n\t"
        "bra PF_WAIT;\n\t"
        "PF_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
// D[tmem] = A[smem] . B[smem], int8 x int8 -> int32, no accumulation (every quant block gets its own TMEM columns)
__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (as vision_kernels.cuh: version 1, layout type 2, SBO = 8 rows * 128 B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::i8: D = S32 (c_format 2), A = B = signed 8-bit (format 1), both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_i8(int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void tmem_ld8i(uint32_t taddr, int (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint4 ldg_nc16(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__host__ __device__ __forceinline__ int bitrev5(int v) { return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4); }
// valid blocks of tile j of class l: blocks b = l + 32 i, i = 4 j .. 4 j + 3, b < nb
__host__ __device__ __forceinline__ int tile_blocks(int nb, int l, int j) {
    const int n_l = nb > l ? (nb - l + 31) / 32 : 0;
    const int v = n_l - 4 * j;
    return v < 0 ? 0 : (v > 4 ? 4 : v);
}

// fused output of one result (row r, token t): res = this row, other = the adjacent row of the pair (RoPE / SwiGLU pair adjacent rows);
// the epilogues of k::matvec_kernel
__device__ __forceinline__ void prefill_epilogue(const PrefillArgs &g, int r, int t, int n_past, float res, float other) {
    if (g.epi == EPI_QKV) {
        const int E = g.e_local, part = r / E, rr = r % E, pos = n_past + t;
        if (part == 2) g.vcache[(size_t)pos * E + rr] = __float2half_rn(res);
        else {
            const float2 cs = g.rope[(size_t)pos * g.half_dim + (rr % (2 * g.half_dim)) / 2];
            const float v0 = (rr & 1) ? other : res, v1 = (rr & 1) ? res : other;
            const float o = (rr & 1) ? (v0 * cs.y + v1 * cs.x) : (v0 * cs.x - v1 * cs.y);
            if (part == 0) g.q_out[(size_t)t * E + rr] = o;
            else g.kcache[(size_t)pos * E + rr] = __float2half_rn(o);
        }
    } else if (g.epi == EPI_RESID) {
        const size_t o = (size_t)t * g.out_stride + r;
        g.out[o] = res + g.resid[o];
    } else if (g.epi == EPI_SWIGLU) {  // rows are interleaved: even = gate(ff), odd = up(ff)
        if ((r & 1) == 0) g.out[(size_t)t * g.out_stride + (r >> 1)] = lut_f16(g.tab_silu, res) * other;
    } else {
        g.out[(size_t)t * g.out_stride + r] = res;
    }
}

// grid (rows_pad / 128, ceil(n_tok / 32)); block 320; dynamic smem = 1024 (alignment) + 4 stages + stack + barriers
template <bool Q41>
__global__ void __launch_bounds__(kThreads, 1) prefill_gemm_q4(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                  const __grid_constant__ CUtensorMap tmS, const PrefillArgs g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    float2 *stack = (float2 *)(smem + (size_t)kStages * kStageBytes);
    uint64_t *full = (uint64_t *)((unsigned char *)stack + kStackBytes);
    uint64_t *empty = full + kStages;
    uint64_t *rfull = empty + kStages;
    uint64_t *rempty = rfull + kGroups;
    uint32_t *tmem_slot = (uint32_t *)(rempty + kGroups);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, t_tile = blockIdx.y;
    const int tpc = g.S >> 2;  // tiles per class
    const int jr0 = (int)blockIdx.z * g.jr_per_z, jr1 = jr0 + g.jr_per_z;   // leaves (classes in bit-reversed order) of this CTA's subtree

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmS) : "memory");
        for (int i = 0; i < kStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1 + kEpiWarps); }
        for (int i = 0; i < kGroups; ++i) { mbar_init(&rfull[i], 1); mbar_init(&rempty[i], kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            unsigned cnt = 0;
            for (int jr = jr0; jr < jr1; ++jr) {
                const int l = bitrev5(jr);
                for (int j = 0; j < tpc; ++j) {
                    if (tile_blocks(g.nb, l, j) == 0) continue;
                    const int s = (int)(cnt % kStages); const uint32_t ph = (cnt / kStages) & 1u; ++cnt;
                    mbar_wait(&empty[s], ph ^ 1u);
                    unsigned char *sa = smem + (size_t)s * kStageBytes;
                    mbar_expect_tx(&full[s], (uint32_t)kStageBytes);
                    const int slot0 = l * g.S + 4 * j;
                    tma_load_2d(sa, &tmA, slot0 * 32, m_tile * kRows, &full[s]);
                    tma_load_2d(sa + 16384, &tmB, slot0 * 32, t_tile * kTok, &full[s]);
                    tma_load_2d(sa + 16384 + 4096, &tmS, slot0 * 8, t_tile * kTok, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = umma_idesc_i8(kTok);
        unsigned cnt = 0;
        for (int jr = jr0; jr < jr1; ++jr) {
            const int l = bitrev5(jr);
            for (int j = 0; j < tpc; ++j) {
                const int v = tile_blocks(g.nb, l, j);
                if (v == 0) continue;
                const int s = (int)(cnt % kStages); const uint32_t ph = (cnt / kStages) & 1u;
                const int gq = (int)(cnt % kGroups); const uint32_t gph = (cnt / kGroups) & 1u; ++cnt;
                mbar_wait(&full[s], ph);
                mbar_wait(&rempty[gq], gph ^ 1u);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + (size_t)s * kStageBytes), sb = sa + 16384u;
                    for (int i = 0; i < v; ++i)
                        tc_mma_i8(tmem_base + (uint32_t)(gq * 128 + i * 32), umma_desc_sw128(sa + i * 32), umma_desc_sw128(sb + i * 32), idesc);
                    tc_commit(&rfull[gq]);   // block dots of this tile are in TMEM
                    tc_commit(&empty[s]);    // ... and the tensor core is done reading the stage's A / B
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------- epilogue: scale, accumulate in canonical order, butterfly tree, fused output ----------------
        const int ew = warp - 2, quarter = warp & 3, tq = ew >> 2;   // TMEM lanes 32 * (warp % 4) .. + 31; tokens kTpw * tq .. + kTpw - 1
        const int row_in = quarter * 32 + lane, r = m_tile * kRows + row_in, c0 = tq * kTpw;
        const uint32_t tlane = (uint32_t)(quarter * 32) << 16;
        const __half2 *wrow = g.wsc + (size_t)r * (32 * g.S);
        float accd[kTpw], accm[kTpw], l0d[kTpw], l0m[kTpw];
#pragma unroll
        for (int c = 0; c < kTpw; ++c) { accd[c] = 0.f; accm[c] = 0.f; l0d[c] = 0.f; l0m[c] = 0.f; }
        unsigned cnt = 0;
        for (int jr = jr0; jr < jr1; ++jr) {
            const int l = bitrev5(jr);
            for (int j = 0; j < tpc; ++j) {
                const int v = tile_blocks(g.nb, l, j);
                if (v == 0) continue;
                const int s = (int)(cnt % kStages); const uint32_t ph = (cnt / kStages) & 1u;
                const int gq = (int)(cnt % kGroups); const uint32_t gph = (cnt / kGroups) & 1u; ++cnt;
                const uint4 w4 = ldg_nc16(wrow + l * g.S + 4 * j);   // {d, m} of this row's four blocks (issued before the waits)
                mbar_wait(&full[s], ph);       // activation scales of the tile are in shared memory
                mbar_wait(&rfull[gq], gph);    // block dots are in TMEM
                tc_fence_after();
                const float2 *asc = (const float2 *)(smem + (size_t)s * kStageBytes + 16384 + 4096);  // [token][4] {d, s}
                const unsigned wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i < v) {
                        int isum[kTpw];
                        tmem_ld8i(tmem_base + tlane + (uint32_t)(gq * 128 + i * 32 + c0), isum);
                        const float2 f = __half22float2(*(const __half2 *)&wv[i]);
#pragma unroll
                        for (int c = 0; c < kTpw; ++c) {
                            const float2 ds = asc[(c0 + c) * 4 + i];
                            if (Q41) { accd[c] = fmaf(f.x * ds.x, (float)isum[c], accd[c]); accm[c] = fmaf(f.y, ds.y, accm[c]); }
                            else accd[c] += ((float)isum[c] * f.x) * ds.x;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&rempty[gq]); mbar_arrive(&empty[s]); }
            }
            // class l is complete: fold it into the butterfly tree (post-order walk, jr = leaf number in bit-reversed class order)
            const int jl = jr - jr0;   // leaf number inside this CTA's subtree (jr0 is a multiple of the subtree size)
            if ((jl & 1) == 0) {
#pragma unroll
                for (int c = 0; c < kTpw; ++c) { l0d[c] = accd[c]; l0m[c] = accm[c]; accd[c] = 0.f; accm[c] = 0.f; }
            } else {
                int ones = 1; while (ones < 5 && ((jl >> ones) & 1)) ++ones;  // trailing ones of jl (>= 1)
#pragma unroll
                for (int c = 0; c < kTpw; ++c) {
                    float vd = l0d[c] + accd[c], vm = l0m[c] + accm[c];
                    accd[c] = 0.f; accm[c] = 0.f;
                    for (int lev = 1; lev < ones; ++lev) {
                        const float2 o = stack[((size_t)(lev - 1) * kTok + (c0 + c)) * kRows + row_in];
                        vd = o.x + vd; vm = o.y + vm;
                    }
                    if (jr != jr1 - 1) stack[((size_t)(ones - 1) * kTok + (c0 + c)) * kRows + row_in] = make_float2(vd, vm);
                    else { l0d[c] = vd; l0m[c] = vm; }
                }
            }
        }
        // results: res[c] = d-tree + m-tree (k::dot2_q4: warp_sum(accd) + warp_sum(accm)); fused epilogues of k::matvec_kernel
        const int n_past = g.state->n_past;
#pragma unroll
        for (int c = 0; c < kTpw; ++c) {
            const int t = t_tile * kTok + c0 + c;
            if (gridDim.z > 1) {   // K split: the root of this CTA's subtree; prefill_combine finishes the tree
                if (t < g.n_tok && r < g.rows) g.partial[((size_t)blockIdx.z * g.part_tok + t) * g.part_rows + r] = make_float2(l0d[c], l0m[c]);
                continue;
            }
            const float res = l0d[c] + l0m[c];
            const float other = __shfl_xor_sync(0xffffffffu, res, 1);   // the pair row (RoPE / SwiGLU pair adjacent rows)
            if (t >= g.n_tok || r >= g.rows) continue;
            prefill_epilogue(g, r, t, n_past, res, other);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// K-split GEMMs: fold the subtree roots of the grid.z slices (pairwise, in leaf order = the upper levels of the butterfly tree), then the epilogue.
// One thread per (token, row pair); nz = 2, 4 or 8.
__global__ void __launch_bounds__(256) prefill_combine(const PrefillArgs g, int nz) {
    const int pairs = g.rows >> 1;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)g.n_tok * pairs) return;
    const int t = (int)(idx / pairs), r0 = 2 * (int)(idx % pairs);
    float res[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        float d[8], m[8];
        for (int z = 0; z < nz; ++z) { const float2 p = g.partial[((size_t)z * g.part_tok + t) * g.part_rows + r0 + k]; d[z] = p.x; m[z] = p.y; }
        for (int st = 1; st < nz; st <<= 1)
            for (int z = 0; z < nz; z += 2 * st) { d[z] = d[z] + d[z + st]; m[z] = m[z] + m[z + st]; }
        res[k] = d[0] + m[0];
    }
    const int n_past = g.state->n_past;
    prefill_epilogue(g, r0, t, n_past, res[0], res[1]);
    prefill_epilogue(g, r0 + 1, t, n_past, res[1], res[0]);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// load time: row-packed Q4 rows ([nb x 16 B nibbles][nb x scales], llama.h) -> int8 class-major operand cache + {d, m} in the same order
// ---------------------------------------------------------------------------------------------------------------------------------
__global__ void expand_q4_classmajor(const unsigned char *src, int row_bytes, int rows, int nb, int q41, int S, signed char *q, __half2 *sc) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * nb) return;
    const int r = (int)(idx / nb), b = (int)(idx % nb);
    const unsigned char *row = src + (size_t)r * row_bytes;
    const uint4 nib = ((const uint4 *)row)[b];
    const unsigned w[4] = {nib.x, nib.y, nib.z, nib.w};
    unsigned lo[4], hi[4];
    for (int k4 = 0; k4 < 4; ++k4) {
        lo[k4] = w[k4] & 0x0F0F0F0Fu; hi[k4] = (w[k4] >> 4) & 0x0F0F0F0Fu;
        if (!q41) { lo[k4] = __vsub4(lo[k4], 0x08080808u); hi[k4] = __vsub4(hi[k4], 0x08080808u); }
    }
    const int slot = (b & 31) * S + (b >> 5);
    uint4 *dst = (uint4 *)(q + ((size_t)r * 32 * S + slot) * 32);
    dst[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);   // elements 0..15 (low nibbles)
    dst[1] = make_uint4(hi[0], hi[1], hi[2], hi[3]);   // elements 16..31 (high nibbles)
    const unsigned char *scp = row + (size_t)nb * 16;
    __half2 v;
    if (q41) v = ((const __half2 *)scp)[b];
    else v = __halves2half2(((const __half *)scp)[b], __ushort_as_half((unsigned short)0));
    sc[(size_t)r * 32 * S + slot] = v;
}

// activations of N rows: (optional RMSNorm) + quantise exactly as k::stage_act, written in the class-major order of the operand cache:
// q8 [n_pad][32 * S * 32] int8, scales [n_pad][32 * S] {d, s}.  grid = n rows, 256 threads, dynamic smem = act_bytes(ACT, cols)
template <int ACT>
__global__ void __launch_bounds__(256) stage_rows_classmajor(const float *x, int x_stride, const float *norm_w, int cols, int S, signed char *q8, float2 *sc) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ double red[34];
    const int t = blockIdx.x, nb = cols >> 5;
    stage_act<ACT>(x + (size_t)t * x_stride, norm_w, cols, sm, red);
    __syncthreads();
    const float *d = (const float *)(sm + cols), *s = d + nb;
    for (int i = threadIdx.x; i < nb * 2; i += 256) {  // 16-byte halves of the blocks: lo plane [b], hi plane [b]
        const int b = i >> 1, hf = i & 1;
        const uint4 v = *(const uint4 *)(sm + (hf ? cols / 2 : 0) + b * 16);
        const int slot = (b & 31) * S + (b >> 5);
        *(uint4 *)(q8 + ((size_t)t * 32 * S + slot) * 32 + hf * 16) = v;
    }
    for (int b = threadIdx.x; b < nb; b += 256) sc[(size_t)t * 32 * S + (b & 31) * S + (b >> 5)] = make_float2(d[b], s[b]);
}

}  // namespace pf
}  // namespace mg4

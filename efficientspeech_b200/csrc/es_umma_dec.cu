// Decoder-side fused tensor-core kernel (sm_100a: tcgen05 + TMEM + bulk-async copies),
// warp-specialised and software-pipelined.
//
// One persistent CTA per SM walks 128-frame tiles of [B, T, 128] activations.  Its 16 warps
// form two groups that work on DIFFERENT tiles at the same time:
//
//   producer warps 8..15 (tile i+1)
//       DWCONV  the x tile (+2-frame halo) arrives by ONE bulk async copy (TMA 1-D, mbarrier
//               complete_tx); depthwise conv k=5 + bias in registers (sliding window)
//       GATHER  the length regulator: row t <- fused4[b, upper_bound(cum[b], t)]
//       PLAIN   rows copied as they are (mel head, stand-alone projection)
//       -> split into fp16 hi/lo, written straight into the UMMA canonical K-major no-swizzle
//          operand layout (bank-conflict free); then ONE thread issues 24 x tcgen05.mma
//          (M128 x N x K16, kind::f16: hi*hi + hi*lo + lo*hi, fp32 accumulate) into one of
//          TWO TMEM accumulators and commits to an mbarrier.  The split-fp16 weights (canonical
//          layout prepared at pack time) are bulk-loaded ONCE per CTA and stay in shared memory.
//   epilogue warps 0..7 (tile i)
//       tcgen05.ld 16x256b: each warp owns 16 complete rows in the mma-fragment layout (4 threads
//       per row), releases the accumulator right after the load, then runs bias -> tanh ->
//       LayerNorm [-> + skip -> LayerNorm] [-> zero padded frames] in registers (2 shuffle steps
//       per statistic) and stores 32-byte row segments straight to global memory.  No
//       shared-memory staging, no CTA-wide barrier in steady state.
//
// mbarriers: bar_x (x tile landed), bar_mma[2] (accumulator s full == A operand free),
// bar_tfree[2] (accumulator s drained by all 8 epilogue warps).  Every wait is bounded: a
// timeout raises a device flag instead of hanging the GPU.
//
// HBM traffic per layer is exactly one read of x (+ skip on block-end layers) and one write of
// y: the kernel is HBM-bound by design.
#include "es_common.cuh"
#include "es_kernels.cuh"
#include "es_umma.cuh"

namespace es {
namespace {

using namespace umma;

constexpr int TM = 128;                 // frames per tile (UMMA M)
constexpr int CK = 128;                 // K = input channels
constexpr int DWK = 5;                  // depthwise taps
constexpr int HALO = DWK / 2;
constexpr int XROWS = TM + DWK - 1;     // 132
constexpr int NTHR = 512;               // 16 warps: 0..7 epilogue, 8..15 producer
constexpr int NPROD = 256;
constexpr uint32_t A_LBO = 144;         // 128-byte core matrix + 16 B pad: conflict-free 8-byte lane stores
constexpr uint32_t A_SBO = 16 * A_LBO;  // 2304: one 8-row group = 16 K-chunks
constexpr uint32_t A_PLANE = 16 * A_SBO;            // 36864 bytes per fp16 plane (hi or lo)
constexpr uint32_t XS_BYTES = XROWS * CK * 4;       // 67584

// shared memory map (dynamic, 1024-aligned base)
constexpr uint32_t OFF_XS = 0;
constexpr uint32_t OFF_A = OFF_XS + XS_BYTES;                 // hi plane, then lo plane
constexpr uint32_t OFF_W = OFF_A + 2 * A_PLANE;               // W hi [K/8][N][8], then W lo
constexpr uint32_t W_PLANE_MAX = 128 * CK * 2;                // 32768
constexpr uint32_t OFF_PAR = OFF_W + 2 * W_PLANE_MAX;         // bias, ln g/b, ln2 g/b: 5 x 128 floats
constexpr uint32_t OFF_SRC = OFF_PAR + 5 * 128 * 4;           // gather sources: 128 ints
constexpr uint32_t OFF_BAR = OFF_SRC + 128 * 4;               // 6 mbarriers + tmem base
constexpr uint32_t SMEM_BYTES = OFF_BAR + 64;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

enum { MODE_DWCONV = 0, MODE_GATHER = 1, MODE_PLAIN = 2 };

struct UmmaDecParams {
    int B, T, N;                 // N = output channels (128, or 80 for the mel head)
    int n_src;                   // GATHER: source rows per utterance
    const float* X;              // DWCONV/PLAIN: [B,T,128]; GATHER: fused4 [B,n_src,128]
    const int* cum;              // GATHER
    const int* valid_len;        // GATHER
    const float* dw_w;           // [5][128]
    const float* dw_b;           // [128]
    const void* w_h16;           // canonical split-fp16 weights: [2][16][N][8] halves
    const float* bias;           // [N]
    int act_tanh;
    const float* ln_g; const float* ln_b;       // LayerNorm over N, or null
    const float* res2; const float* ln2_g; const float* ln2_b;   // out = LN2(out + res2), or null
    const int* zero_from;        // [B] rows t >= zero_from[b] zeroed, or null
    float* Y;                    // [B,T,N]
    int* err;                    // device error flag (mbarrier timeout)
};

// tanh(x) = 1 - 2 / (1 + e^{2x}) with e^{2x} = 2^{x * 2 log2(e)}: two MUFU ops (ex2, rcp) and two FMAs.
// The argument arrives pre-scaled (acc * c + bias * c, c = 2 log2 e).  Saturates correctly at
// +-inf (ex2 -> inf -> rcp -> 0 -> 1; ex2 -> 0 -> rcp(1) -> -1); absolute error ~1.2e-7.
constexpr float kTanhScale = 2.8853900817779268f;
__device__ __forceinline__ float tanh_from_scaled(float arg) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(arg));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return fmaf(-2.f, r, 1.f);
}

__device__ __forceinline__ uint2 pack_half4(float a, float b, float c, float d) {
    const __half2 p = __floats2half2_rn(a, b), q = __floats2half2_rn(c, d);
    return make_uint2(*reinterpret_cast<const uint32_t*>(&p), *reinterpret_cast<const uint32_t*>(&q));
}

// writes 4 consecutive channels (4*lane .. 4*lane+3) of tile row `row` as split fp16 into the A planes
__device__ __forceinline__ void store_a4(uint8_t* a_hi, int row, int lane, float4 v) {
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const float2 b0 = __half22float2(h0), b1 = __half22float2(h1);
    const uint2 hi = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    const uint2 lo = pack_half4(v.x - b0.x, v.y - b0.y, v.z - b1.x, v.w - b1.y);
    const uint32_t off = (uint32_t)(lane >> 1) * A_LBO + (uint32_t)(row >> 3) * A_SBO + (uint32_t)(row & 7) * 16u +
                         (uint32_t)(lane & 1) * 8u;
    *reinterpret_cast<uint2*>(a_hi + off) = hi;
    *reinterpret_cast<uint2*>(a_hi + A_PLANE + off) = lo;
}

// LayerNorm of two rows held in the 16x256b fragment layout: v[4j+2i+b] = row i, column 8j+2*t4+b.
// The 4 threads of a quad (t4 = 0..3) hold one row: 2 xor-shuffles per statistic.
template <int NJ>
__device__ __forceinline__ void fragment_layernorm(float (&v)[64], const float* __restrict__ g,
                                                   const float* __restrict__ be, int t4, float inv_n) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) { s0 += v[4 * j] + v[4 * j + 1]; s1 += v[4 * j + 2] + v[4 * j + 3]; }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float m0 = s0 * inv_n, m1 = s1 * inv_n;
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        v[4 * j] -= m0; v[4 * j + 1] -= m0; v[4 * j + 2] -= m1; v[4 * j + 3] -= m1;
        q0 = fmaf(v[4 * j], v[4 * j], q0); q0 = fmaf(v[4 * j + 1], v[4 * j + 1], q0);
        q1 = fmaf(v[4 * j + 2], v[4 * j + 2], q1); q1 = fmaf(v[4 * j + 3], v[4 * j + 3], q1);
    }
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
    q0 += __shfl_xor_sync(0xffffffffu, q0, 2); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
    const float r0 = rsqrtf(q0 * inv_n + kLnEps), r1 = rsqrtf(q1 * inv_n + kLnEps);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const float2 gg = *reinterpret_cast<const float2*>(g + 8 * j + 2 * t4);
        const float2 bb = *reinterpret_cast<const float2*>(be + 8 * j + 2 * t4);
        v[4 * j] = fmaf(v[4 * j] * r0, gg.x, bb.x);
        v[4 * j + 1] = fmaf(v[4 * j + 1] * r0, gg.y, bb.y);
        v[4 * j + 2] = fmaf(v[4 * j + 2] * r1, gg.x, bb.x);
        v[4 * j + 3] = fmaf(v[4 * j + 3] * r1, gg.y, bb.y);
    }
}

template <int MODE>
__global__ void __launch_bounds__(NTHR, 1)
umma_dec_kernel(const UmmaDecParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* Xs = reinterpret_cast<float*>(smem + OFF_XS);
    uint8_t* a_hi = smem + OFF_A;
    float* par = reinterpret_cast<float*>(smem + OFF_PAR);
    int* srcs = reinterpret_cast<int*>(smem + OFF_SRC);
    const uint32_t bar_x = smem_u32(smem + OFF_BAR), bar_w = bar_x + 8;
    const uint32_t bar_mma = bar_x + 16;      // [2]
    const uint32_t bar_tfree = bar_x + 32;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 48);

    const int N = p.N;
    const int tiles_per_utt = (p.T + TM - 1) / TM;
    const int n_tiles = p.B * tiles_per_utt;
    const uint32_t w_plane = (uint32_t)N * CK * 2u;

    // ---- one-time setup ---------------------------------------------------------------------
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);      // two 128-column fp32 accumulators
    if (tid == 0) {
        mbar_init(bar_x, 1);
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        mbar_init(bar_mma + 8, 1);
        mbar_init(bar_tfree, 8);
        mbar_init(bar_tfree + 8, 8);
        fence_mbar_init();
    }
    for (int i = tid; i < 128; i += NTHR) {
        par[i] = (i < N) ? __ldg(p.bias + i) * (p.act_tanh ? kTanhScale : 1.f) : 0.f;   // pre-scaled for tanh
        par[128 + i] = (p.ln_g && i < N) ? __ldg(p.ln_g + i) : 0.f;
        par[256 + i] = (p.ln_g && i < N) ? __ldg(p.ln_b + i) : 0.f;
        par[384 + i] = (p.ln2_g && i < N) ? __ldg(p.ln2_g + i) : 0.f;
        par[512 + i] = (p.ln2_g && i < N) ? __ldg(p.ln2_b + i) : 0.f;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    bool failed = false;

    if (warp >= 8) {
        // =========================================================================== producers
        const int pw = warp - 8, ptid = tid - NPROD;
        const bool leader = (ptid == 0);

        auto issue_x = [&](int tile) {     // rows [t0-HALO, t0+TM+HALO) clipped to the utterance, one bulk copy
            const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TM;
            const int lo = max(t0 - HALO, 0), hi = min(t0 + TM + HALO, p.T);
            const uint32_t bytes = (uint32_t)(hi - lo) * CK * 4u;
            mbar_arrive_expect_tx(bar_x, bytes);
            bulk_g2s(smem_u32(Xs) + (uint32_t)(lo - (t0 - HALO)) * CK * 4u,
                     p.X + ((size_t)b * p.T + lo) * CK, bytes, bar_x);
        };
        if (leader) {
            mbar_arrive_expect_tx(bar_w, 2 * w_plane);
            bulk_g2s(smem_u32(smem + OFF_W), p.w_h16, w_plane, bar_w);
            bulk_g2s(smem_u32(smem + OFF_W) + w_plane, reinterpret_cast<const uint8_t*>(p.w_h16) + w_plane, w_plane, bar_w);
            if (MODE == MODE_DWCONV && (int)blockIdx.x < n_tiles) issue_x(blockIdx.x);
        }
        // per-lane depthwise taps for channels 4*lane..4*lane+3 (persistent in registers)
        float4 wdw[DWK], bdw;
        if (MODE == MODE_DWCONV) {
#pragma unroll
            for (int t = 0; t < DWK; ++t) wdw[t] = __ldg(reinterpret_cast<const float4*>(p.dw_w + t * CK) + lane);
            bdw = __ldg(reinterpret_cast<const float4*>(p.dw_b) + lane);
        }
        const uint32_t idesc = make_idesc_f16(TM, N);
        const uint32_t lbo_b = (uint32_t)N * 16u;

        int i = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
            const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TM;
            const int rows_valid = min(TM, p.T - t0);
            const int s = i & 1, u = i >> 1;
            // A operand free?  (MMA of the previous tile has read it)
            if (i > 0 && !mbar_wait(bar_mma + 8 * ((i - 1) & 1), ((i - 1) >> 1) & 1)) failed = true;

            if (MODE == MODE_DWCONV) {
                // zero the halo / tail rows the bulk copy does not cover (utterance boundaries only)
                const int lo = max(t0 - HALO, 0), hi = min(t0 + TM + HALO, p.T);
                const int head = lo - (t0 - HALO), tail0 = hi - (t0 - HALO);
                if (head > 0 || tail0 < XROWS) {
                    for (int k = ptid; k < (head + XROWS - tail0) * (CK / 4); k += NPROD) {
                        int r = k / (CK / 4);
                        const int c4 = k - r * (CK / 4);
                        if (r >= head) r = tail0 + (r - head);
                        reinterpret_cast<float4*>(Xs + r * CK)[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                if (!mbar_wait(bar_x, i & 1)) failed = true;
                named_bar_sync(1, NPROD);              // zero fill visible to every producer warp
#pragma unroll 1
                for (int g = 0; g < 2; ++g) {          // warp pw -> output rows 16pw..16pw+15, 8 at a time
                    const int r0 = pw * 16 + g * 8;
                    float4 win[12];
#pragma unroll
                    for (int k = 0; k < 12; ++k) win[k] = reinterpret_cast<const float4*>(Xs + (r0 + k) * CK)[lane];
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        float4 o = bdw;
#pragma unroll
                        for (int t = 0; t < DWK; ++t) {
                            o.x = fmaf(wdw[t].x, win[r + t].x, o.x);
                            o.y = fmaf(wdw[t].y, win[r + t].y, o.y);
                            o.z = fmaf(wdw[t].z, win[r + t].z, o.z);
                            o.w = fmaf(wdw[t].w, win[r + t].w, o.w);
                        }
                        store_a4(a_hi, r0 + r, lane, o);
                    }
                }
            } else {
                if (MODE == MODE_GATHER) {
                    if (ptid < TM) {
                        const int t = t0 + ptid;
                        int sidx = -1;
                        if (t < p.T && t < p.valid_len[b]) {
                            const int* c = p.cum + (size_t)b * p.n_src;
                            int lo = 0, hi = p.n_src;
                            while (lo < hi) {
                                const int mid = (lo + hi) >> 1;
                                if (__ldg(c + mid) > t) hi = mid; else lo = mid + 1;
                            }
                            sidx = lo < p.n_src ? lo : -1;
                        }
                        srcs[ptid] = sidx;
                    }
                    named_bar_sync(1, NPROD);
                }
#pragma unroll 4
                for (int r = 0; r < 16; ++r) {
                    const int row = pw * 16 + r;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (MODE == MODE_GATHER) {
                        const int sidx = srcs[row];
                        if (sidx >= 0) v = __ldg(reinterpret_cast<const float4*>(p.X + ((size_t)b * p.n_src + sidx) * CK) + lane);
                    } else if (row < rows_valid) {
                        v = __ldg(reinterpret_cast<const float4*>(p.X + ((size_t)b * p.T + t0 + row) * CK) + lane);
                    }
                    store_a4(a_hi, row, lane, v);
                }
            }
            fence_proxy_async_smem();                  // generic-proxy stores -> visible to the tensor core
            tc_fence_before_sync();
            named_bar_sync(1, NPROD);

            if (leader) {
                tc_fence_after_sync();
                // x of the next tile streams in while this tile's GEMM / epilogue run (Xs is free:
                // every producer warp passed the barrier above after its last read)
                if (MODE == MODE_DWCONV && tile + (int)gridDim.x < n_tiles) issue_x(tile + gridDim.x);
                if (i == 0 && !mbar_wait(bar_w, 0)) failed = true;
                // accumulator s drained by the epilogue warps (its previous use was tile i-2)?
                if (u > 0 && !mbar_wait(bar_tfree + 8 * s, (u - 1) & 1)) failed = true;
                tc_fence_after_sync();
                const uint32_t a0 = smem_u32(a_hi), w0 = smem_u32(smem + OFF_W);
                const uint32_t acc = tmem + (uint32_t)(s * 128);
#pragma unroll 1
                for (int k = 0; k < CK / 16; ++k) {
                    const uint64_t dah = make_smem_desc(a0 + (uint32_t)(2 * k) * A_LBO, A_LBO, A_SBO);
                    const uint64_t dal = make_smem_desc(a0 + A_PLANE + (uint32_t)(2 * k) * A_LBO, A_LBO, A_SBO);
                    const uint64_t dbh = make_smem_desc(w0 + (uint32_t)(2 * k) * lbo_b, lbo_b, 128u);
                    const uint64_t dbl = make_smem_desc(w0 + w_plane + (uint32_t)(2 * k) * lbo_b, lbo_b, 128u);
                    mma_f16_ss(acc, dah, dbh, idesc, k > 0 ? 1u : 0u);
                    mma_f16_ss(acc, dah, dbl, idesc, 1u);
                    mma_f16_ss(acc, dal, dbh, idesc, 1u);
                }
                mma_commit(bar_mma + 8 * s);
            }
        }
    } else {
        // =========================================================================== epilogue
        const int q = warp & 3, half = warp >> 2;             // TMEM lane quarter, 16-row half
        const int rbase = q * 32 + half * 16;
        const int t4 = lane & 3, tr = lane >> 2;
        const float inv_n = 1.f / (float)N;
        const int nj = N >> 3;                                // 8-column groups

        int i = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
            const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TM;
            const int rows_valid = min(TM, p.T - t0);
            const int s = i & 1, u = i >> 1;
            const int row0 = rbase + tr, row1 = row0 + 8;
            const size_t g0 = (size_t)b * p.T + t0 + row0, g1 = g0 + 8;
            const bool ok0 = row0 < rows_valid, ok1 = row1 < rows_valid;

            if (p.res2) {   // pull this warp's 16 skip rows (8 KB) towards L2 while the GEMM runs
                const int pr = rbase + (lane >> 1);
                if (pr < rows_valid) {
                    const float* sp = p.res2 + ((size_t)b * p.T + t0 + pr) * N + (lane & 1) * 64;
                    prefetch_l2(sp);
                    prefetch_l2(sp + 32);
                }
            }
            if (!mbar_wait(bar_mma + 8 * s, u & 1)) failed = true;
            tc_fence_after_sync();
            uint32_t r[64];
            tmem_ld_16x256b_x16(tmem + (uint32_t)(s * 128) + ((uint32_t)rbase << 16), r);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tfree + 8 * s);     // accumulator drained: the GEMM of tile i+2 may start

            float v[64];
            if (p.act_tanh) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 bb = *reinterpret_cast<const float2*>(par + 8 * j + 2 * t4);
                    v[4 * j] = tanh_from_scaled(fmaf(__uint_as_float(r[4 * j]), kTanhScale, bb.x));
                    v[4 * j + 1] = tanh_from_scaled(fmaf(__uint_as_float(r[4 * j + 1]), kTanhScale, bb.y));
                    v[4 * j + 2] = tanh_from_scaled(fmaf(__uint_as_float(r[4 * j + 2]), kTanhScale, bb.x));
                    v[4 * j + 3] = tanh_from_scaled(fmaf(__uint_as_float(r[4 * j + 3]), kTanhScale, bb.y));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 bb = *reinterpret_cast<const float2*>(par + 8 * j + 2 * t4);
                    v[4 * j] = __uint_as_float(r[4 * j]) + bb.x;
                    v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb.y;
                    v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb.x;
                    v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb.y;
                }
            }
            if (p.ln_g) fragment_layernorm<16>(v, par + 128, par + 256, t4, inv_n);   // LayerNorm only with N == 128
            if (p.res2) {
                const float* s0 = p.res2 + g0 * N + 2 * t4;
                const float* s1 = p.res2 + g1 * N + 2 * t4;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float2 a = ok0 ? __ldg(reinterpret_cast<const float2*>(s0 + 8 * j)) : make_float2(0.f, 0.f);
                    const float2 c = ok1 ? __ldg(reinterpret_cast<const float2*>(s1 + 8 * j)) : make_float2(0.f, 0.f);
                    v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += c.x; v[4 * j + 3] += c.y;
                }
                fragment_layernorm<16>(v, par + 384, par + 512, t4, inv_n);
            }
            const int zero_from = p.zero_from ? p.zero_from[b] : 0x7fffffff;
            const bool z0 = (t0 + row0) >= zero_from, z1 = (t0 + row1) >= zero_from;
            float* y0 = p.Y + g0 * N + 2 * t4;
            float* y1 = p.Y + g1 * N + 2 * t4;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (j < nj) {
                    if (ok0) *reinterpret_cast<float2*>(y0 + 8 * j) = z0 ? make_float2(0.f, 0.f) : make_float2(v[4 * j], v[4 * j + 1]);
                    if (ok1) *reinterpret_cast<float2*>(y1 + 8 * j) = z1 ? make_float2(0.f, 0.f) : make_float2(v[4 * j + 2], v[4 * j + 3]);
                }
            }
        }
    }

    if (failed) atomicExch(p.err, 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

int* g_err_flag = nullptr;

template <int MODE>
int launch_mode(const UmmaDecParams& p, int grid, cudaStream_t s) {
    static bool attr_set = false;
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(umma_dec_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    umma_dec_kernel<MODE><<<grid, NTHR, SMEM_BYTES, s>>>(p);
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace

bool umma_dec_supported(int C, int dw_k, int N) {
    return C == CK && dw_k == DWK && (N == 128 || N == 80);
}

// mode: 0 depthwise layer, 1 gather + projection, 2 plain (mel head)
int launch_umma_dec(int mode, int B, int T, int N, int n_src, const float* X, const int* cum,
                    const int* valid_len, const float* dw_w, const float* dw_b, const void* w_h16,
                    const float* bias, int act_tanh, const float* ln_g, const float* ln_b,
                    const float* res2, const float* ln2_g, const float* ln2_b, const int* zero_from,
                    float* Y, cudaStream_t s) {
    ES_CHECK(w_h16 && X && Y && bias, "null tensor");
    ES_CHECK(N % 16 == 0 && N >= 32 && N <= 128, "N must be a multiple of 16 in [32,128]");
    ES_CHECK(!(ln_g || res2) || N == 128, "LayerNorm epilogue needs N == 128");
    if (!g_err_flag) {
        ES_CUDA(cudaMalloc(&g_err_flag, sizeof(int)));
        ES_CUDA(cudaMemset(g_err_flag, 0, sizeof(int)));
    }
    static int n_sm = 0;
    if (!n_sm) {
        int dev = 0;
        ES_CUDA(cudaGetDevice(&dev));
        ES_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    UmmaDecParams p;
    p.B = B; p.T = T; p.N = N; p.n_src = n_src; p.X = X; p.cum = cum; p.valid_len = valid_len;
    p.dw_w = dw_w; p.dw_b = dw_b; p.w_h16 = w_h16; p.bias = bias; p.act_tanh = act_tanh;
    p.ln_g = ln_g; p.ln_b = ln_b; p.res2 = res2; p.ln2_g = ln2_g; p.ln2_b = ln2_b;
    p.zero_from = zero_from; p.Y = Y; p.err = g_err_flag;
    const int n_tiles = B * ((T + TM - 1) / TM);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    switch (mode) {
        case MODE_DWCONV: return launch_mode<MODE_DWCONV>(p, grid, s);
        case MODE_GATHER: return launch_mode<MODE_GATHER>(p, grid, s);
        default: return launch_mode<MODE_PLAIN>(p, grid, s);
    }
}

// Reads (and clears) the device-side mbarrier-timeout flag; synchronises the stream.
int umma_dec_check_errors(cudaStream_t s) {
    if (!g_err_flag) return 0;
    int h = 0;
    ES_CUDA(cudaMemcpyAsync(&h, g_err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    ES_CUDA(cudaStreamSynchronize(s));
    if (h) {
        ES_CUDA(cudaMemsetAsync(g_err_flag, 0, sizeof(int), s));
        ES_CHECK(false, "a tcgen05 kernel timed out waiting on an mbarrier");
    }
    return 0;
}

}  // namespace es

// Decoder-side fused tensor-core kernel (sm_100a, tcgen05 + TMEM + bulk-async copies).
//
// One persistent CTA per SM walks 128-frame tiles of [B, T, 128] activations:
//
//   prologue   DWCONV  x tile (+2-frame halo) arrives by ONE bulk async copy (TMA 1-D, mbarrier
//                      complete_tx); depthwise conv k=5 + bias in registers (sliding window)
//              GATHER  the length regulator: row t <- fused4[b, upper_bound(cum[b], t)]
//              PLAIN   rows copied as they are (mel head)
//              -> split into fp16 hi/lo and written straight into the UMMA canonical K-major
//                 no-swizzle operand layout (bank-conflict free, see es_umma.cuh)
//   GEMM       24 x tcgen05.mma (M128 x N x K16, kind::f16; hi*hi + hi*lo + lo*hi) issued by one
//              thread, fp32 accumulator in TMEM, completion signalled on an mbarrier by
//              tcgen05.commit; the weights (split fp16, canonical layout prepared at pack time)
//              are loaded ONCE per CTA and stay resident in shared memory
//   epilogue A tcgen05.ld (thread = tile row) -> + bias -> tanh -> XOR-swizzled smem staging
//   epilogue B warp per row: LayerNorm (shuffle reductions) [-> + skip -> LayerNorm] [-> zero
//              padded frames] -> 512-byte coalesced global stores
//
// The next tile's x is in flight (bulk copy) while the current tile runs its GEMM + epilogue.
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
constexpr int NTHR = 512;               // 16 warps
constexpr uint32_t A_LBO = 144;         // 128-byte core matrix + 16 B pad: conflict-free 8-byte lane stores
constexpr uint32_t A_SBO = 16 * A_LBO;  // 2304: one 8-row group = 16 K-chunks
constexpr uint32_t A_PLANE = 16 * A_SBO;            // 36864 bytes per fp16 plane (hi or lo)
constexpr uint32_t XS_BYTES = XROWS * CK * 4;       // 67584
constexpr uint32_t STG_ROW = 512;                   // staging row stride (bytes)

// shared memory map (dynamic, 1024-aligned base)
constexpr uint32_t OFF_XS = 0;
constexpr uint32_t OFF_A = OFF_XS + XS_BYTES;                 // hi plane, then lo plane; reused as staging
constexpr uint32_t OFF_W = OFF_A + 2 * A_PLANE;               // W hi [K/8][N][8], then W lo
constexpr uint32_t W_PLANE_MAX = 128 * CK * 2;                // 32768
constexpr uint32_t OFF_PAR = OFF_W + 2 * W_PLANE_MAX;         // bias, ln g/b, ln2 g/b: 5 x 128 floats
constexpr uint32_t OFF_SRC = OFF_PAR + 5 * 128 * 4;           // gather sources: 128 ints
constexpr uint32_t OFF_BAR = OFF_SRC + 128 * 4;               // 3 mbarriers + tmem base
constexpr uint32_t SMEM_BYTES = OFF_BAR + 64;
static_assert(2 * A_PLANE >= TM * STG_ROW, "staging must fit in the A operand region");
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

template <int MODE>
__global__ void __launch_bounds__(NTHR, 1)
umma_dec_kernel(const UmmaDecParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* Xs = reinterpret_cast<float*>(smem + OFF_XS);
    uint8_t* a_hi = smem + OFF_A;
    uint8_t* stg = smem + OFF_A;
    float* par = reinterpret_cast<float*>(smem + OFF_PAR);
    int* srcs = reinterpret_cast<int*>(smem + OFF_SRC);
    const uint32_t bar_x = smem_u32(smem + OFF_BAR), bar_w = bar_x + 8, bar_m = bar_x + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 24);

    const int N = p.N;
    const int tiles_per_utt = (p.T + TM - 1) / TM;
    const int n_tiles = p.B * tiles_per_utt;
    const uint32_t w_plane = (uint32_t)N * CK * 2u;

    // ---- one-time setup ---------------------------------------------------------------------
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 128);
    if (tid == 0) {
        mbar_init(bar_x, 1);
        mbar_init(bar_w, 1);
        mbar_init(bar_m, 1);
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

    // x-tile loader: rows [t0-HALO, t0+TM+HALO) clipped to the utterance, one bulk copy
    auto issue_x = [&](int tile) {
        const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TM;
        const int lo = max(t0 - HALO, 0), hi = min(t0 + TM + HALO, p.T);
        const uint32_t bytes = (uint32_t)(hi - lo) * CK * 4u;
        mbar_arrive_expect_tx(bar_x, bytes);
        bulk_g2s(smem_u32(Xs) + (uint32_t)(lo - (t0 - HALO)) * CK * 4u,
                 p.X + ((size_t)b * p.T + lo) * CK, bytes, bar_x);
    };

    if (tid == 0) {
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

    bool failed = false;
    if (!mbar_wait(bar_w, 0)) failed = true;

    const uint32_t idesc = make_idesc_f16(TM, N);
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, phase ^= 1) {
        const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TM;
        const int rows_valid = min(TM, p.T - t0);

        // ---------------------------------------------------------------- prologue -> A operand
        if (MODE == MODE_DWCONV) {
            // zero the halo / tail rows the bulk copy does not cover (utterance boundaries only)
            const int lo = max(t0 - HALO, 0), hi = min(t0 + TM + HALO, p.T);
            const int head = lo - (t0 - HALO), tail0 = hi - (t0 - HALO);
            if (head > 0 || tail0 < XROWS) {
                for (int i = tid; i < (head + XROWS - tail0) * (CK / 4); i += NTHR) {
                    int r = i / (CK / 4);
                    const int c4 = i - r * (CK / 4);
                    if (r >= head) r = tail0 + (r - head);
                    reinterpret_cast<float4*>(Xs + r * CK)[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            if (!mbar_wait(bar_x, phase)) failed = true;
            __syncthreads();                       // zero fill visible to every warp
            // warp g -> output rows 8g..8g+7 (one 8-row core-matrix group); lane -> 4 channels
            const int r0 = warp * 8;
            float4 win[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) win[i] = reinterpret_cast<const float4*>(Xs + (r0 + i) * CK)[lane];
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
        } else {
            if (MODE == MODE_GATHER) {
                if (tid < TM) {
                    const int t = t0 + tid;
                    int s = -1;
                    if (t < p.T && t < p.valid_len[b]) {
                        const int* c = p.cum + (size_t)b * p.n_src;
                        int lo = 0, hi = p.n_src;
                        while (lo < hi) {
                            const int mid = (lo + hi) >> 1;
                            if (__ldg(c + mid) > t) hi = mid; else lo = mid + 1;
                        }
                        s = lo < p.n_src ? lo : -1;
                    }
                    srcs[tid] = s;
                }
                __syncthreads();
            }
            const int r0 = warp * 8;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int row = r0 + r;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (MODE == MODE_GATHER) {
                    const int s = srcs[row];
                    if (s >= 0) v = __ldg(reinterpret_cast<const float4*>(p.X + ((size_t)b * p.n_src + s) * CK) + lane);
                } else if (row < rows_valid) {
                    v = __ldg(reinterpret_cast<const float4*>(p.X + ((size_t)b * p.T + t0 + row) * CK) + lane);
                }
                store_a4(a_hi, row, lane, v);
            }
        }
        fence_proxy_async_smem();                  // generic-proxy stores -> visible to the tensor core
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();

        // ---------------------------------------------------------------- GEMM (one thread issues)
        if (tid == 0) {
            const uint32_t a0 = smem_u32(a_hi), w0 = smem_u32(smem + OFF_W);
            const uint32_t lbo_b = (uint32_t)N * 16u;
#pragma unroll 1
            for (int s = 0; s < CK / 16; ++s) {
                const uint64_t dah = make_smem_desc(a0 + (uint32_t)(2 * s) * A_LBO, A_LBO, A_SBO);
                const uint64_t dal = make_smem_desc(a0 + A_PLANE + (uint32_t)(2 * s) * A_LBO, A_LBO, A_SBO);
                const uint64_t dbh = make_smem_desc(w0 + (uint32_t)(2 * s) * lbo_b, lbo_b, 128u);
                const uint64_t dbl = make_smem_desc(w0 + w_plane + (uint32_t)(2 * s) * lbo_b, lbo_b, 128u);
                mma_f16_ss(tmem, dah, dbh, idesc, s > 0 ? 1u : 0u);
                mma_f16_ss(tmem, dah, dbl, idesc, 1u);
                mma_f16_ss(tmem, dal, dbh, idesc, 1u);
            }
            mma_commit(bar_m);
            // x of the next tile streams in while the GEMM and the epilogue run (Xs is free: every
            // warp passed the barrier above after its last read)
            if (MODE == MODE_DWCONV && tile + (int)gridDim.x < n_tiles) issue_x(tile + gridDim.x);
        }
        if (!mbar_wait(bar_m, phase)) failed = true;
        tc_fence_after_sync();

        // ---------------------------------------------------------------- epilogue A: TMEM -> staging
        {
            const int q = warp & 3, cq = warp >> 2;          // TMEM lane quarter (rows), column quarter
            const int row = q * 32 + lane;
            if (cq * 32 < N) {
                uint32_t r[32];
                tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cq * 32), r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bb = *reinterpret_cast<const float4*>(par + cq * 32 + j * 4);
                    float4 v;
                    if (p.act_tanh) {
                        v.x = tanh_from_scaled(fmaf(__uint_as_float(r[4 * j]), kTanhScale, bb.x));
                        v.y = tanh_from_scaled(fmaf(__uint_as_float(r[4 * j + 1]), kTanhScale, bb.y));
                        v.z = tanh_from_scaled(fmaf(__uint_as_float(r[4 * j + 2]), kTanhScale, bb.z));
                        v.w = tanh_from_scaled(fmaf(__uint_as_float(r[4 * j + 3]), kTanhScale, bb.w));
                    } else {
                        v = make_float4(__uint_as_float(r[4 * j]) + bb.x, __uint_as_float(r[4 * j + 1]) + bb.y,
                                        __uint_as_float(r[4 * j + 2]) + bb.z, __uint_as_float(r[4 * j + 3]) + bb.w);
                    }
                    const int chunk = cq * 8 + j;
                    *reinterpret_cast<float4*>(stg + (uint32_t)row * STG_ROW + (uint32_t)((chunk ^ (row & 7)) * 16)) = v;
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();

        // ---------------------------------------------------------------- epilogue B
        // 4 rows per warp step, 8 lanes per row, 16 channels per lane (16-byte chunks part + 8j):
        // LayerNorm statistics need 3 shuffle steps for 4 rows at once, every global access is a
        // full 128-byte line per 8-lane group.
        {
            const int rr = lane >> 3, part = lane & 7;
            const int nl = N >> 2;                            // 16-byte chunks per row
            const float inv_n = 1.f / (float)N;
            const int zero_from = p.zero_from ? p.zero_from[b] : 0x7fffffff;
#pragma unroll 1
            for (int rbase = warp * 4; rbase < TM; rbase += (NTHR / 32) * 4) {
                const int row = rbase + rr;
                const bool rvalid = row < rows_valid;
                const size_t grow = (size_t)b * p.T + t0 + row;
                float4 v[4];
                bool cv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    cv[j] = (part + 8 * j) < nl;
                    v[j] = cv[j] ? *reinterpret_cast<const float4*>(stg + (uint32_t)row * STG_ROW +
                                                                    (uint32_t)((8 * j + (part ^ (row & 7))) * 16))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (p.ln_g) {
                    float sm = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) sm += (v[j].x + v[j].y) + (v[j].z + v[j].w);
                    sm += __shfl_xor_sync(0xffffffffu, sm, 1);
                    sm += __shfl_xor_sync(0xffffffffu, sm, 2);
                    sm += __shfl_xor_sync(0xffffffffu, sm, 4);
                    const float mean = sm * inv_n;
                    float q = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
                        if (!cv[j]) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        q = fmaf(v[j].x, v[j].x, q); q = fmaf(v[j].y, v[j].y, q);
                        q = fmaf(v[j].z, v[j].z, q); q = fmaf(v[j].w, v[j].w, q);
                    }
                    q += __shfl_xor_sync(0xffffffffu, q, 1);
                    q += __shfl_xor_sync(0xffffffffu, q, 2);
                    q += __shfl_xor_sync(0xffffffffu, q, 4);
                    const float rstd = rsqrtf(q * inv_n + kLnEps);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 g = *reinterpret_cast<const float4*>(par + 128 + (part + 8 * j) * 4);
                        const float4 be = *reinterpret_cast<const float4*>(par + 256 + (part + 8 * j) * 4);
                        v[j] = make_float4(fmaf(v[j].x * rstd, g.x, be.x), fmaf(v[j].y * rstd, g.y, be.y),
                                           fmaf(v[j].z * rstd, g.z, be.z), fmaf(v[j].w * rstd, g.w, be.w));
                    }
                }
                if (p.res2) {
                    float sm = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (cv[j] && rvalid) {
                            const float4 sk = __ldg(reinterpret_cast<const float4*>(p.res2 + grow * N) + part + 8 * j);
                            v[j].x += sk.x; v[j].y += sk.y; v[j].z += sk.z; v[j].w += sk.w;
                        }
                        sm += (v[j].x + v[j].y) + (v[j].z + v[j].w);
                    }
                    sm += __shfl_xor_sync(0xffffffffu, sm, 1);
                    sm += __shfl_xor_sync(0xffffffffu, sm, 2);
                    sm += __shfl_xor_sync(0xffffffffu, sm, 4);
                    const float mean = sm * inv_n;
                    float q = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
                        if (!cv[j]) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        q = fmaf(v[j].x, v[j].x, q); q = fmaf(v[j].y, v[j].y, q);
                        q = fmaf(v[j].z, v[j].z, q); q = fmaf(v[j].w, v[j].w, q);
                    }
                    q += __shfl_xor_sync(0xffffffffu, q, 1);
                    q += __shfl_xor_sync(0xffffffffu, q, 2);
                    q += __shfl_xor_sync(0xffffffffu, q, 4);
                    const float rstd = rsqrtf(q * inv_n + kLnEps);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 g = *reinterpret_cast<const float4*>(par + 384 + (part + 8 * j) * 4);
                        const float4 be = *reinterpret_cast<const float4*>(par + 512 + (part + 8 * j) * 4);
                        v[j] = make_float4(fmaf(v[j].x * rstd, g.x, be.x), fmaf(v[j].y * rstd, g.y, be.y),
                                           fmaf(v[j].z * rstd, g.z, be.z), fmaf(v[j].w * rstd, g.w, be.w));
                    }
                }
                if (rvalid) {
                    const bool zero = (t0 + row) >= zero_from;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (cv[j]) reinterpret_cast<float4*>(p.Y + grow * N)[part + 8 * j] =
                            zero ? make_float4(0.f, 0.f, 0.f, 0.f) : v[j];
                }
            }
        }
        __syncthreads();                           // staging (== A region) free for the next tile
    }

    if (failed && lane == 0) atomicExch(p.err, 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
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

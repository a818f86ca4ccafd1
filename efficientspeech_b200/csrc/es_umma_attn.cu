// Unmasked multi-head self-attention on the tensor cores (sm_100a: tcgen05 + TMEM).
//
// One work item = one (utterance, head) with n <= 128 positions and head width C in {32, 64}
// (blocks.py:43-66: every head is full width C, scale = (C // H)^-0.5, softmax over ALL n keys --
// the reference builds a mask and never applies it).  A CTA of 128 threads (thread = query row =
// TMEM lane) processes items in a persistent loop; two CTAs share an SM so that one item's
// load / softmax phases overlap the other's GEMMs:
//
//   1. Q, K rows -> split fp16 hi/lo -> UMMA canonical K-major row-panel layout; V is stored
//      TRANSPOSED ([C rows][n keys], K-major) so that it can be the B operand of the second GEMM
//   2. S = Q K^T        3 x tcgen05.mma per K step (hi*hi + hi*lo + lo*hi), M128 x N128, fp32 in TMEM
//   3. softmax          thread-per-row straight out of TMEM (tcgen05.ld 32x32b): max, sum of
//                       ex2((s - max) * scale * log2 e), then P = e / sum is split to fp16 hi/lo and
//                       written over the dead Q/K tiles as the A operand of the second GEMM
//   4. O = P V          24 x tcgen05.mma (K = 128 keys), M128 x N=C
//   5. O rows -> global [B, n, H*C] at column h*C
//
// Keys >= n do not exist (they are not "masked"): their S columns are skipped and their P is 0.
#include "es_common.cuh"
#include "es_kernels.cuh"
#include "es_umma.cuh"

namespace es {
namespace {

using namespace umma;

constexpr int AT_M = 128;                        // query rows per item (TMEM lanes)
constexpr uint32_t PANEL = AT_M * 16;            // 2048: one K panel (8 elements) of 128 rows
constexpr uint32_t P_PLANE = 16 * PANEL;         // 128 keys -> 16 panels = 32768
// region R0 holds {Q hi, Q lo, K hi, K lo} (<= 64 KB) and is later overwritten by {P hi, P lo} (64 KB)
constexpr uint32_t OFF_R0 = 0;
constexpr uint32_t OFF_VT = OFF_R0 + 2 * P_PLANE;             // V^T hi, lo: [16 panels][C rows][8] each
constexpr uint32_t VT_PLANE_MAX = 16 * 64 * 16;               // 16384
constexpr uint32_t OFF_BAR = OFF_VT + 2 * VT_PLANE_MAX;
constexpr uint32_t ATT_SMEM = OFF_BAR + 64;                   // 98368 B -> two CTAs per SM
static_assert(2 * ATT_SMEM <= 227 * 1024, "two CTAs per SM");

struct AttnParams {
    const float* qkv;     // [B, n, 3*H*C], channel order [q|k|v][head][c]
    float* out;           // [B, n, H*C]
    int B, n, C, H;
    int kp;               // wide kernel: key count padded to 64 / 128
    float scale_log2e;    // (C // H)^-0.5 * log2(e)
    int* err;
};

__global__ void __launch_bounds__(128, 2)
umma_attention_kernel(const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int n = p.n, C = p.C, H = p.H;
    const uint32_t bar = smem_u32(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 16);
    const uint32_t qk_plane = (uint32_t)(C >> 3) * PANEL;      // bytes of one Q or K plane
    uint8_t* q_hi = smem + OFF_R0;
    uint8_t* k_hi = q_hi + 2 * qk_plane;
    uint8_t* p_hi = smem + OFF_R0;
    uint8_t* vt_hi = smem + OFF_VT;
    const uint32_t vt_lbo = (uint32_t)C * 16u;                 // V^T: one K panel (8 keys) of C rows
    const uint32_t vt_plane = 16u * vt_lbo;

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);       // S: columns 0..127, O: columns 128..128+C
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    pdl_launch_dependents();      // the next kernel may start its prologue
    pdl_wait();                   // the QKV projection's output is complete and visible from here on
    const bool elected_warp = (warp == 0);
    bool failed = false;
    uint32_t phase = 0;

    const size_t ldq = (size_t)3 * H * C;
    const int items = p.B * H;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = item / H, h = item - b * H;
        const float* base = p.qkv + (size_t)b * n * ldq;
        const int r = tid;                                     // this thread's row (query / key index)
        // ------------------------------------------------------------ 1. stage Q, K, V^T
        {
            const bool live = r < n;
            const float4* qp = reinterpret_cast<const float4*>(base + (size_t)r * ldq + h * C);
            const float4* kp = reinterpret_cast<const float4*>(base + (size_t)r * ldq + (H + h) * C);
            const float4* vp = reinterpret_cast<const float4*>(base + (size_t)r * ldq + (2 * H + h) * C);
            for (int pc = 0; pc < (C >> 3); ++pc) {            // one K panel (8 channels) at a time
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0, c0 = a0, c1 = a0;
                if (live) {
                    a0 = __ldg(qp + 2 * pc); a1 = __ldg(qp + 2 * pc + 1);
                    b0 = __ldg(kp + 2 * pc); b1 = __ldg(kp + 2 * pc + 1);
                    c0 = __ldg(vp + 2 * pc); c1 = __ldg(vp + 2 * pc + 1);
                }
                const float qa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float ka[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                uint4 hi, lo;
                split8(qa, hi, lo);
                *reinterpret_cast<uint4*>(q_hi + (uint32_t)pc * PANEL + (uint32_t)r * 16u) = hi;
                *reinterpret_cast<uint4*>(q_hi + qk_plane + (uint32_t)pc * PANEL + (uint32_t)r * 16u) = lo;
                split8(ka, hi, lo);
                *reinterpret_cast<uint4*>(k_hi + (uint32_t)pc * PANEL + (uint32_t)r * 16u) = hi;
                *reinterpret_cast<uint4*>(k_hi + qk_plane + (uint32_t)pc * PANEL + (uint32_t)r * 16u) = lo;
                // V^T[c][key r]: K panel r/8, row c, element r%8
                const float va[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const __half vh = __float2half_rn(va[e]);
                    const __half vl = __float2half_rn(va[e] - __half2float(vh));
                    const uint32_t off = (uint32_t)(r >> 3) * vt_lbo + (uint32_t)(pc * 8 + e) * 16u + (uint32_t)(r & 7) * 2u;
                    *reinterpret_cast<__half*>(vt_hi + off) = vh;
                    *reinterpret_cast<__half*>(vt_hi + vt_plane + off) = vl;
                }
            }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
        // ------------------------------------------------------------ 2. S = Q K^T
        if (elected_warp) {
            const bool elected = elect_one();
            const uint32_t idesc = make_idesc_f16(AT_M, 128);
            const uint32_t q0 = smem_u32(q_hi), k0 = smem_u32(k_hi);
            for (int ks = 0; ks < (C >> 4); ++ks) {
                const uint32_t o = (uint32_t)(2 * ks) * PANEL;
                const uint64_t dqh = make_smem_desc(q0 + o, PANEL, 128u), dql = make_smem_desc(q0 + qk_plane + o, PANEL, 128u);
                const uint64_t dkh = make_smem_desc(k0 + o, PANEL, 128u), dkl = make_smem_desc(k0 + qk_plane + o, PANEL, 128u);
                if (elected) {
                    mma_f16_ss(tmem, dqh, dkh, idesc, ks > 0 ? 1u : 0u);
                    mma_f16_ss(tmem, dqh, dkl, idesc, 1u);
                    mma_f16_ss(tmem, dql, dkh, idesc, 1u);
                }
            }
            if (elected) mma_commit(bar);
            __syncwarp();
        }
        if (!mbar_wait(bar, phase)) failed = true;
        phase ^= 1;
        tc_fence_after_sync();
        // ------------------------------------------------------------ 3. softmax (thread = row), P -> smem
        {
            const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
            float mx = -INFINITY;
            for (int c0 = 0; c0 < n; c0 += 32) {
                uint32_t rr[32];
                tmem_ld32(trow + (uint32_t)c0, rr);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < n) mx = fmaxf(mx, __uint_as_float(rr[j]));
            }
            const float sc = p.scale_log2e;                    // > 0, so max(scale * s) = scale * max(s)
            const float nm = -mx * sc;
            float sum = 0.f;
            for (int c0 = 0; c0 < n; c0 += 32) {
                uint32_t rr[32];
                tmem_ld32(trow + (uint32_t)c0, rr);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float e;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(__uint_as_float(rr[j]), sc, nm)));
                    if (c0 + j < n) sum += e;
                }
            }
            const float inv = 1.f / sum;
            for (int c0 = 0; c0 < AT_M; c0 += 32) {
                uint32_t rr[32];
                if (c0 < n) {
                    tmem_ld32(trow + (uint32_t)c0, rr);
                    tmem_ld_wait();
                }
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    float pv[8];
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8) {
                        const int col = c0 + j8 * 8 + e8;
                        float e = 0.f;
                        if (col < n) {
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(__uint_as_float(rr[j8 * 8 + e8]), sc, nm)));
                            e *= inv;
                        }
                        pv[e8] = e;
                    }
                    uint4 hi, lo;
                    split8(pv, hi, lo);
                    const uint32_t off = (uint32_t)((c0 >> 3) + j8) * PANEL + (uint32_t)tid * 16u;
                    *reinterpret_cast<uint4*>(p_hi + off) = hi;          // Q/K tiles are dead: S is complete
                    *reinterpret_cast<uint4*>(p_hi + P_PLANE + off) = lo;
                }
            }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
        // ------------------------------------------------------------ 4. O = P V
        if (elected_warp) {
            const bool elected = elect_one();
            const uint32_t idesc = make_idesc_f16(AT_M, C);
            const uint32_t p0 = smem_u32(p_hi), v0 = smem_u32(vt_hi);
            for (int ks = 0; ks < AT_M / 16; ++ks) {
                const uint64_t dph = make_smem_desc(p0 + (uint32_t)(2 * ks) * PANEL, PANEL, 128u);
                const uint64_t dpl = make_smem_desc(p0 + P_PLANE + (uint32_t)(2 * ks) * PANEL, PANEL, 128u);
                const uint64_t dvh = make_smem_desc(v0 + (uint32_t)(2 * ks) * vt_lbo, vt_lbo, 128u);
                const uint64_t dvl = make_smem_desc(v0 + vt_plane + (uint32_t)(2 * ks) * vt_lbo, vt_lbo, 128u);
                if (elected) {
                    mma_f16_ss(tmem + 128, dph, dvh, idesc, ks > 0 ? 1u : 0u);
                    mma_f16_ss(tmem + 128, dph, dvl, idesc, 1u);
                    mma_f16_ss(tmem + 128, dpl, dvh, idesc, 1u);
                }
            }
            if (elected) mma_commit(bar);
            __syncwarp();
        }
        if (!mbar_wait(bar, phase)) failed = true;
        phase ^= 1;
        tc_fence_after_sync();
        // ------------------------------------------------------------ 5. O rows -> global
        {
            const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16) + 128u;
            float* orow = p.out + ((size_t)b * n + r) * ((size_t)H * C) + h * C;
            for (int c0 = 0; c0 < C; c0 += 32) {
                uint32_t rr[32];
                tmem_ld32(trow + (uint32_t)c0, rr);
                tmem_ld_wait();
                if (r < n) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1]),
                                                                                __uint_as_float(rr[j + 2]), __uint_as_float(rr[j + 3]));
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();                                       // smem tiles and TMEM reusable by the next item
        tc_fence_after_sync();
    }

    if (failed) atomicExch(p.err, 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------------------------
// Wide heads (base: C = 128 with n <= 128, C = 256 with n <= 64; blocks.py gives every head the FULL width C).
// Same five phases, with the head width streamed: Q / K go through the 64 KB region 64 channels at a time (each chunk is
// exactly the narrow kernel's C = 64 staging) and S accumulates across chunks; V^T [C rows][KP keys] is staged next to
// it chunk by chunk.  KP = 64 or 128 is the key count padded to the MMA's K granularity for P V, and bounds C * KP to
// the 64 KB V^T region.  O (N = C <= 256 columns) overwrites S in TMEM once the softmax has moved P to shared memory.
// One CTA per SM (128 KB of shared memory).
constexpr uint32_t ATW_SMEM = OFF_R0 + 2 * P_PLANE + 65536 + 64;            // 131136 B
constexpr uint32_t ATW_OFF_VT = OFF_R0 + 2 * P_PLANE;
constexpr uint32_t ATW_OFF_BAR = ATW_OFF_VT + 65536;

__global__ void __launch_bounds__(128, 1)
umma_attention_wide_kernel(const AttnParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int n = p.n, C = p.C, H = p.H, KP = p.kp;
    const uint32_t bar = smem_u32(smem + ATW_OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ATW_OFF_BAR + 16);
    constexpr uint32_t qk_plane = 8u * PANEL;                  // 64 channels of 128 rows: 16 KB
    uint8_t* q_hi = smem + OFF_R0;
    uint8_t* k_hi = q_hi + 2 * qk_plane;
    uint8_t* p_hi = smem + OFF_R0;
    uint8_t* vt_hi = smem + ATW_OFF_VT;
    const uint32_t vt_lbo = (uint32_t)C * 16u;                 // V^T: one K panel (8 keys) of C rows
    const uint32_t vt_plane = (uint32_t)(KP >> 3) * vt_lbo;    // <= 32 KB
    const uint32_t p_plane = (uint32_t)(KP >> 3) * PANEL;      // P: KP keys of 128 query rows

    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);       // S: columns 0..KP-1, then O: columns 0..C-1
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();
    const bool elected_warp = (warp == 0);
    bool failed = false;
    uint32_t phase = 0;

    const size_t ldq = (size_t)3 * H * C;
    const int items = p.B * H;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = item / H, h = item - b * H;
        const float* base = p.qkv + (size_t)b * n * ldq;
        const int r = tid;
        const bool live = r < n;
        // ------------------------------------------------------------ 1+2. stream the head width: stage 64 channels, S += Q K^T
        for (int ch = 0; ch < (C >> 6); ++ch) {
            const float4* qp = reinterpret_cast<const float4*>(base + (size_t)r * ldq + h * C + ch * 64);
            const float4* kp = reinterpret_cast<const float4*>(base + (size_t)r * ldq + (H + h) * C + ch * 64);
            const float4* vp = reinterpret_cast<const float4*>(base + (size_t)r * ldq + (2 * H + h) * C + ch * 64);
            for (int pc = 0; pc < 8; ++pc) {
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0, c0 = a0, c1 = a0;
                if (live) {
                    a0 = __ldg(qp + 2 * pc); a1 = __ldg(qp + 2 * pc + 1);
                    b0 = __ldg(kp + 2 * pc); b1 = __ldg(kp + 2 * pc + 1);
                    c0 = __ldg(vp + 2 * pc); c1 = __ldg(vp + 2 * pc + 1);
                }
                const float qa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float ka[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                uint4 hi, lo;
                split8(qa, hi, lo);
                *reinterpret_cast<uint4*>(q_hi + (uint32_t)pc * PANEL + (uint32_t)r * 16u) = hi;
                *reinterpret_cast<uint4*>(q_hi + qk_plane + (uint32_t)pc * PANEL + (uint32_t)r * 16u) = lo;
                split8(ka, hi, lo);
                *reinterpret_cast<uint4*>(k_hi + (uint32_t)pc * PANEL + (uint32_t)r * 16u) = hi;
                *reinterpret_cast<uint4*>(k_hi + qk_plane + (uint32_t)pc * PANEL + (uint32_t)r * 16u) = lo;
                if (r < KP) {                                  // V^T[c][key r]: key panel r/8, row c, element r%8 (keys >= n are zero)
                    const float va[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const __half vh = __float2half_rn(va[e]);
                        const __half vl = __float2half_rn(va[e] - __half2float(vh));
                        const uint32_t off = (uint32_t)(r >> 3) * vt_lbo + (uint32_t)(ch * 64 + pc * 8 + e) * 16u + (uint32_t)(r & 7) * 2u;
                        *reinterpret_cast<__half*>(vt_hi + off) = vh;
                        *reinterpret_cast<__half*>(vt_hi + vt_plane + off) = vl;
                    }
                }
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncthreads();
            tc_fence_after_sync();
            if (elected_warp) {
                const bool elected = elect_one();
                const uint32_t idesc = make_idesc_f16(AT_M, KP);
                const uint32_t q0 = smem_u32(q_hi), k0 = smem_u32(k_hi);
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t o = (uint32_t)(2 * ks) * PANEL;
                    const uint64_t dqh = make_smem_desc(q0 + o, PANEL, 128u), dql = make_smem_desc(q0 + qk_plane + o, PANEL, 128u);
                    const uint64_t dkh = make_smem_desc(k0 + o, PANEL, 128u), dkl = make_smem_desc(k0 + qk_plane + o, PANEL, 128u);
                    if (elected) {
                        mma_f16_ss(tmem, dqh, dkh, idesc, (ch > 0 || ks > 0) ? 1u : 0u);
                        mma_f16_ss(tmem, dqh, dkl, idesc, 1u);
                        mma_f16_ss(tmem, dql, dkh, idesc, 1u);
                    }
                }
                if (elected) mma_commit(bar);
                __syncwarp();
            }
            if (!mbar_wait(bar, phase)) failed = true;         // the chunk buffers are free again; after the last chunk S is complete
            phase ^= 1;
            tc_fence_after_sync();
        }
        // ------------------------------------------------------------ 3. softmax (thread = row), P -> smem
        {
            const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
            float mx = -INFINITY;
            for (int c0 = 0; c0 < n; c0 += 32) {
                uint32_t rr[32];
                tmem_ld32(trow + (uint32_t)c0, rr);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (c0 + j < n) mx = fmaxf(mx, __uint_as_float(rr[j]));
            }
            const float sc = p.scale_log2e;
            const float nm = -mx * sc;
            float sum = 0.f;
            for (int c0 = 0; c0 < n; c0 += 32) {
                uint32_t rr[32];
                tmem_ld32(trow + (uint32_t)c0, rr);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float e;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(__uint_as_float(rr[j]), sc, nm)));
                    if (c0 + j < n) sum += e;
                }
            }
            const float inv = 1.f / sum;
            for (int c0 = 0; c0 < KP; c0 += 32) {
                uint32_t rr[32];
                if (c0 < n) {
                    tmem_ld32(trow + (uint32_t)c0, rr);
                    tmem_ld_wait();
                }
#pragma unroll
                for (int j8 = 0; j8 < 4; ++j8) {
                    float pv[8];
#pragma unroll
                    for (int e8 = 0; e8 < 8; ++e8) {
                        const int col = c0 + j8 * 8 + e8;
                        float e = 0.f;
                        if (col < n) {
                            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(__uint_as_float(rr[j8 * 8 + e8]), sc, nm)));
                            e *= inv;
                        }
                        pv[e8] = e;
                    }
                    uint4 hi, lo;
                    split8(pv, hi, lo);
                    const uint32_t off = (uint32_t)((c0 >> 3) + j8) * PANEL + (uint32_t)tid * 16u;
                    *reinterpret_cast<uint4*>(p_hi + off) = hi;          // the Q / K chunk is dead: S is complete
                    *reinterpret_cast<uint4*>(p_hi + p_plane + off) = lo;
                }
            }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();                                       // every row's S has been read: O may overwrite it
        tc_fence_after_sync();
        // ------------------------------------------------------------ 4. O = P V  (K = KP keys, N = C)
        if (elected_warp) {
            const bool elected = elect_one();
            const uint32_t idesc = make_idesc_f16(AT_M, C);
            const uint32_t p0 = smem_u32(p_hi), v0 = smem_u32(vt_hi);
            for (int ks = 0; ks < (KP >> 4); ++ks) {
                const uint64_t dph = make_smem_desc(p0 + (uint32_t)(2 * ks) * PANEL, PANEL, 128u);
                const uint64_t dpl = make_smem_desc(p0 + p_plane + (uint32_t)(2 * ks) * PANEL, PANEL, 128u);
                const uint64_t dvh = make_smem_desc(v0 + (uint32_t)(2 * ks) * vt_lbo, vt_lbo, 128u);
                const uint64_t dvl = make_smem_desc(v0 + vt_plane + (uint32_t)(2 * ks) * vt_lbo, vt_lbo, 128u);
                if (elected) {
                    mma_f16_ss(tmem, dph, dvh, idesc, ks > 0 ? 1u : 0u);
                    mma_f16_ss(tmem, dph, dvl, idesc, 1u);
                    mma_f16_ss(tmem, dpl, dvh, idesc, 1u);
                }
            }
            if (elected) mma_commit(bar);
            __syncwarp();
        }
        if (!mbar_wait(bar, phase)) failed = true;
        phase ^= 1;
        tc_fence_after_sync();
        // ------------------------------------------------------------ 5. O rows -> global
        {
            const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
            float* orow = p.out + ((size_t)b * n + r) * ((size_t)H * C) + h * C;
            for (int c0 = 0; c0 < C; c0 += 32) {
                uint32_t rr[32];
                tmem_ld32(trow + (uint32_t)c0, rr);
                tmem_ld_wait();
                if (live) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(__uint_as_float(rr[j]), __uint_as_float(rr[j + 1]),
                                                                                __uint_as_float(rr[j + 2]), __uint_as_float(rr[j + 3]));
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
    }

    if (failed) atomicExch(p.err, 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

// Returns -1 when the shape is outside the tensor-core kernel's envelope (caller uses the SIMT kernel).
int launch_umma_attention(const float* qkv, float* out, int B, int n, int C, int H, float scale, cudaStream_t s) {
    if (n > AT_M || n < 1) return -1;
    const bool wide = (C == 128 || C == 256);
    const int KP = n <= 64 ? 64 : 128;                         // keys padded for the wide kernel
    if (!wide && C != 32 && C != 64) return -1;
    if (wide && C * KP > 16384) return -1;                     // V^T must fit its 64 KB region (C = 256 needs n <= 64)
    int* err_flag = umma_err_flag();
    ES_CHECK(err_flag, "cannot allocate the device error flag");
    static PerDeviceSlot<int> n_sm_once; int& n_sm = n_sm_once.get();
    if (!n_sm) {
        int dev = 0;
        ES_CUDA(cudaGetDevice(&dev));
        ES_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    static PerDeviceSlot<bool> attr_once; bool& attr_set = attr_once.get();   // function attributes are per device
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(umma_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        ES_CUDA(cudaFuncSetAttribute(umma_attention_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATW_SMEM));
        attr_set = true;
    }
    AttnParams p;
    p.qkv = qkv; p.out = out; p.B = B; p.n = n; p.C = C; p.H = H; p.kp = KP;
    p.scale_log2e = scale * 1.4426950408889634f;
    p.err = err_flag;
    const int items = B * H;
    if (wide) {
        ES_CUDA(launch_pdl(umma_attention_wide_kernel, items < n_sm ? items : n_sm, 128, ATW_SMEM, s, p));
        ES_LAUNCH_OK();
        return 0;
    }
    const int grid = items < 2 * n_sm ? items : 2 * n_sm;
    ES_CUDA(launch_pdl(umma_attention_kernel, grid, 128, ATT_SMEM, s, p));
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace es

// Phoneme-side dense layers on the tensor cores (sm_100a: tcgen05 + TMEM), generic over K / N /
// conv taps.  Same warp-specialised pipeline as the decoder kernel (es_umma_dec.cu):
//
//   producer warps 8..11  load the 64 (+taps-1 halo) input rows of a tile from global memory
//                         (the phoneme-side tensors are L2 resident), split them into fp16 hi/lo
//                         and store them in the UMMA canonical K-major no-swizzle layout with
//                         ROW PANELS: addr(row, k) = (k/8)*LBO + row*16 + (k%8)*2.  Rows of a K
//                         panel are 16 bytes apart, so the operand of conv tap tau is THE SAME
//                         buffer with its descriptor start address advanced by tau*16 bytes -- a
//                         dense Conv1d(k=3) is three accumulating GEMMs over one staged tile.
//   issue warp 12         for every tap and K step: 3 x tcgen05.mma (M64 x N<=128 x K16, split
//                         fp16) per 128-column chunk of N; commit to an mbarrier.  Weights
//                         ([tap][hi,lo][K/8][N][8], prepared at pack time) are bulk-loaded once
//                         per CTA and stay resident.
//   epilogue warps 0..7   tcgen05.ld 16x256b per 128-column chunk; bias, boundary-aware tap bias,
//                         ReLU / exact-erf GELU, scalar head (quad shuffle), residual, Fuse scatter,
//                         LayerNorm, second activation, padding mask, 8-byte stores -- all in registers.
//
// Used for: QKV / attention-output projections, the folded MixFFN conv and mlp2, the block-1
// merge conv (stride 2, 1 tap), both GEMMs of the folded Fuse (d <= 64) and both convs of the three variance predictors whenever
// K <= 128, N <= 384 and the weights fit in shared memory; everything else stays on the fp32
// SIMT kernels of es_rowgemm.cu.
#include <stdlib.h>

#include "es_common.cuh"
#include "es_kernels.cuh"
#include "es_umma.cuh"

namespace es {
namespace {

using namespace umma;

constexpr int TM = 64;                   // rows per tile (UMMA M)
constexpr int NTHR = 416;                // 13 warps: 0..7 epilogue, 8..11 producer, 12 issue
constexpr int NPROD = 128;
constexpr int KMAX = 128;
constexpr int MAXTAPS = 3;
constexpr int ROWS_A = TM + MAXTAPS - 1; // 66 staged rows
constexpr uint32_t A_LBO = ROWS_A * 16;  // 1056: one K panel (8 channels) of all staged rows
constexpr uint32_t A_PLANE = (KMAX / 8) * A_LBO;        // 16896 bytes (hi or lo)
constexpr uint32_t A_STAGE = 2 * A_PLANE;               // 33792
constexpr uint32_t W_MAX = 100 * 1024;                  // resident weights budget
constexpr int NPAR = 384;                               // longest parameter vector

constexpr uint32_t OFF_A = 0;                           // 2 stages
constexpr uint32_t OFF_W = OFF_A + 2 * A_STAGE;         // 67584
constexpr uint32_t OFF_PAR = OFF_W + W_MAX;             // bias[384] tapb[3][384] dotw[128] lng[128] lnb[128]
constexpr uint32_t PAR_FLOATS = NPAR + MAXTAPS * NPAR + 3 * 128;
constexpr uint32_t OFF_BAR = OFF_PAR + PAR_FLOATS * 4;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 128;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct UrgParams {
    RowGemmParams g[3];          // up to 3 independent problems of identical geometry (the three predictors)
    const void* w_h16[3];
    int nprob;
    int* err;
};

// NJ = number of valid 8-column groups in every accumulator chunk (N = 8*NJ for N <= 128, 16 otherwise):
// the epilogue loops are specialised on it, so a 32-channel layer does a quarter of the work.
template <int NJ>
__global__ void __launch_bounds__(NTHR, 1)
umma_rowgemm_kernel(const UrgParams up) {
    // CTA -> problem: consecutive CTAs serve different problems; each keeps its own weights resident
    const int prob = blockIdx.x % up.nprob;
    const int cta = blockIdx.x / up.nprob, ncta = gridDim.x / up.nprob;
    const RowGemmParams& p = up.g[prob];
    const void* w_h16 = up.w_h16[prob];
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* par = reinterpret_cast<float*>(smem + OFF_PAR);
    float* s_bias = par;
    float* s_tapb = par + NPAR;
    float* s_dotw = par + NPAR + MAXTAPS * NPAR;
    float* s_lng = s_dotw + 128;
    float* s_lnb = s_lng + 128;
    const uint32_t bar_w = smem_u32(smem + OFF_BAR);
    const uint32_t bar_aready = bar_w + 8;                 // [2] A stage written by the 4 producer warps
    const uint32_t bar_mma = bar_w + 24;                   // [2] accumulator g full == A stage g free
    const uint32_t bar_tfree = bar_w + 40;                 // [2] accumulator g drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 64);

    const int K = p.K, N = p.Nout, taps = p.taps;
    const int tiles_per_utt = (p.n_out + TM - 1) / TM;
    const int n_tiles = p.B * tiles_per_utt;
    const uint32_t w_plane = (uint32_t)N * K * 2u;         // one fp16 plane of one tap
    const int nchunks = (N + 127) >> 7;                    // 128-column accumulator chunks
    const int rows_a = (p.stride == 1) ? TM + taps - 1 : TM;

    // ---- one-time setup ---------------------------------------------------------------------
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    if (tid == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_aready, 4);
        mbar_init(bar_aready + 8, 4);
        mbar_init(bar_mma, 1);
        mbar_init(bar_mma + 8, 1);
        mbar_init(bar_tfree, 4);
        mbar_init(bar_tfree + 8, 4);
        fence_mbar_init();
        // the weights do not depend on the predecessor kernel: their bulk load starts before pdl_wait()
        mbar_arrive_expect_tx(bar_w, (uint32_t)taps * 2u * w_plane);
        for (int k = 0; k < taps * 2; ++k)
            bulk_g2s(smem_u32(smem + OFF_W) + (uint32_t)k * w_plane,
                     reinterpret_cast<const uint8_t*>(w_h16) + (size_t)k * w_plane, w_plane, bar_w);
    }
    for (int i = tid; i < NPAR; i += NTHR) {
        s_bias[i] = (p.bias && i < N) ? __ldg(p.bias + i) : 0.f;
        for (int t = 0; t < MAXTAPS; ++t)
            s_tapb[t * NPAR + i] = (p.tap_bias && t < taps && i < N) ? __ldg(p.tap_bias + t * p.ldw + i) : 0.f;
        if (i < 128) {
            s_dotw[i] = (p.dot_w && i < N) ? __ldg(p.dot_w + i) : 0.f;
            s_lng[i] = (p.ln_g && i < N) ? __ldg(p.ln_g + i) : 0.f;
            s_lnb[i] = (p.ln_g && i < N) ? __ldg(p.ln_b + i) : 0.f;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    bool failed = false;
    pdl_launch_dependents();      // the next kernel may start its prologue
    pdl_wait();                   // the previous kernel's output is complete and visible from here on

    if (warp == 12) {
        // =========================================================================== issue warp
        const bool elected = elect_one();
        const uint32_t idesc_full = make_idesc_f16(TM, N >= 128 ? 128 : N);
        const uint32_t idesc_last = make_idesc_f16(TM, (N & 127) ? (N & 127) : 128);
        const uint32_t lbo_b = (uint32_t)N * 16u;
        if (!mbar_wait(bar_w, 0)) failed = true;
        __syncwarp();
        int i = 0;
        for (int tile = cta; tile < n_tiles; tile += ncta, ++i) {
            const int g = i & 1, u = i >> 1;
            if (!mbar_wait(bar_aready + 8 * g, u & 1)) failed = true;
            if (u > 0 && !mbar_wait(bar_tfree + 8 * g, (u - 1) & 1)) failed = true;
            tc_fence_after_sync();
            const uint32_t a_base = smem_u32(smem + OFF_A) + (uint32_t)g * A_STAGE;
            const uint32_t w_base = smem_u32(smem + OFF_W);
            for (int c = 0; c < nchunks; ++c) {
                const uint32_t acc = tmem + ((uint32_t)(16 * g) << 16) + (uint32_t)(c * 128);
                const uint32_t idesc = (c == nchunks - 1) ? idesc_last : idesc_full;
                uint32_t first = 1;
                for (int t = 0; t < taps; ++t) {
                    for (int ks = 0; ks < (K >> 4); ++ks) {
                        // A: stage g, rows shifted by tap t (16 bytes per row inside a K panel)
                        const uint32_t a_off = (uint32_t)t * 16u + (uint32_t)(2 * ks) * A_LBO;
                        const uint64_t dah = make_smem_desc(a_base + a_off, A_LBO, 128u);
                        const uint64_t dal = make_smem_desc(a_base + A_PLANE + a_off, A_LBO, 128u);
                        // B: tap t, rows (output channels) of chunk c, K step ks
                        const uint32_t b_off = (uint32_t)(t * 2) * w_plane + (uint32_t)(c * 128) * 16u + (uint32_t)(2 * ks) * lbo_b;
                        const uint64_t dbh = make_smem_desc(w_base + b_off, lbo_b, 128u);
                        const uint64_t dbl = make_smem_desc(w_base + w_plane + b_off, lbo_b, 128u);
                        if (elected) {
                            mma_f16_ss(acc, dah, dbh, idesc, first ? 0u : 1u);
                            mma_f16_ss(acc, dah, dbl, idesc, 1u);
                            mma_f16_ss(acc, dal, dbh, idesc, 1u);
                        }
                        first = 0;
                    }
                }
            }
            if (elected) mma_commit(bar_mma + 8 * g);
            __syncwarp();
        }
    } else if (warp >= 8) {
        // =========================================================================== producers
        const int ptid = tid - 256;
        const int K4 = K >> 2;
        int i = 0;
        for (int tile = cta; tile < n_tiles; tile += ncta, ++i) {
            const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TM;
            const int g = i & 1, u = i >> 1;
            uint8_t* a_hi = smem + OFF_A + (uint32_t)g * A_STAGE;
            const float* Ab = p.A + (size_t)b * p.n_in * p.lda;
            const int first_in = t0 * p.stride - p.pad;
            const int total = rows_a * K4;
            // stage g free?  (its previous GEMM, tile i-2, has completed)
            bool waited = (u == 0);
            for (int base = 0; base < total; base += NPROD * 8) {
                float4 v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {                 // up to 8 loads in flight per thread
                    const int idx = base + q * NPROD + ptid;
                    v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (idx < total) {
                        const int r = idx / K4, c4 = idx - r * K4;
                        const int t_in = first_in + r * p.stride;
                        if (t_in >= 0 && t_in < p.n_in) v[q] = __ldg(reinterpret_cast<const float4*>(Ab + (size_t)t_in * p.lda) + c4);
                    }
                }
                if (!waited) {
                    if (!mbar_wait(bar_mma + 8 * g, (u - 1) & 1)) failed = true;
                    waited = true;
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int idx = base + q * NPROD + ptid;
                    if (idx < total) {
                        const int r = idx / K4, c4 = idx - r * K4;
                        uint2 hi, lo;
                        split4(v[q], hi, lo);
                        const uint32_t off = (uint32_t)(c4 >> 1) * A_LBO + (uint32_t)r * 16u + (uint32_t)(c4 & 1) * 8u;
                        *reinterpret_cast<uint2*>(a_hi + off) = hi;
                        *reinterpret_cast<uint2*>(a_hi + A_PLANE + off) = lo;
                    }
                }
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_aready + 8 * g);
        }
    } else {
        // =========================================================================== epilogue
        const int q = warp & 3, g = warp >> 2;
        const int rbase = q * 16;
        const int t4 = lane & 3, tr = lane >> 2;
        const float inv_n = 1.f / (float)N;
        const bool full_epi = (N <= 128);

        int i = g;
        for (int tile = cta + g * ncta; tile < n_tiles; tile += 2 * ncta, i += 2) {
            const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TM;
            const int rows_valid = min(TM, p.n_out - t0);
            const int u = i >> 1;
            const int row0 = rbase + tr, row1 = row0 + 8;
            const int tt0 = t0 + row0, tt1 = t0 + row1;
            const size_t g0 = (size_t)b * p.n_out + tt0, g1 = g0 + 8;
            const bool ok0 = row0 < rows_valid, ok1 = row1 < rows_valid;
            if (!mbar_wait(bar_mma + 8 * g, u & 1)) failed = true;
            tc_fence_after_sync();

            for (int c = 0; c < nchunks; ++c) {
                uint32_t r[64];
                tmem_ld_16x256b_x16(tmem + ((uint32_t)(32 * q + 16 * g) << 16) + (uint32_t)(c * 128), r);
                tmem_ld_wait();
                if (c == nchunks - 1) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tfree + 8 * g);
                }
                const int cbase = c * 128;
                float v[64];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float2 bb = *reinterpret_cast<const float2*>(s_bias + cbase + 8 * j + 2 * t4);
                    v[4 * j] = __uint_as_float(r[4 * j]) + bb.x;
                    v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb.y;
                    v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb.x;
                    v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb.y;
                }
                if (p.tap_bias) {
                    for (int t = 0; t < taps; ++t) {
                        const int ti0 = tt0 * p.stride + t - p.pad, ti1 = tt1 * p.stride + t - p.pad;
                        const float m0 = (ti0 >= 0 && ti0 < p.n_in) ? 1.f : 0.f, m1 = (ti1 >= 0 && ti1 < p.n_in) ? 1.f : 0.f;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            const float2 tb = *reinterpret_cast<const float2*>(s_tapb + t * NPAR + cbase + 8 * j + 2 * t4);
                            v[4 * j] = fmaf(m0, tb.x, v[4 * j]); v[4 * j + 1] = fmaf(m0, tb.y, v[4 * j + 1]);
                            v[4 * j + 2] = fmaf(m1, tb.x, v[4 * j + 2]); v[4 * j + 3] = fmaf(m1, tb.y, v[4 * j + 3]);
                        }
                    }
                }
                if (p.act1 == ACT_RELU) {
#pragma unroll
                    for (int k = 0; k < 4 * NJ; ++k) v[k] = fmaxf(v[k], 0.f);
                } else if (p.act1 == ACT_GELU) {               // exact erf GELU (blocks.py:19)
#pragma unroll
                    for (int k = 0; k < 4 * NJ; ++k) v[k] = gelu_erf_f(v[k]);
                }
                if (full_epi) {
                    if (p.dot_out) {
                        float d0 = 0.f, d1 = 0.f;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            const float2 w = *reinterpret_cast<const float2*>(s_dotw + 8 * j + 2 * t4);
                            d0 = fmaf(v[4 * j], w.x, d0); d0 = fmaf(v[4 * j + 1], w.y, d0);
                            d1 = fmaf(v[4 * j + 2], w.x, d1); d1 = fmaf(v[4 * j + 3], w.y, d1);
                        }
                        d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
                        d0 += __shfl_xor_sync(0xffffffffu, d0, 2); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
                        const float db = __ldg(p.dot_b);
                        d0 += db; d1 += db;
                        if (p.dot_relu) { d0 = fmaxf(d0, 0.f); d1 = fmaxf(d1, 0.f); }
                        if (t4 == 0) {
                            if (ok0) p.dot_out[g0] = d0;
                            if (ok1) p.dot_out[g1] = d1;
                        }
                    }
                    if (p.res1) {
                        const float* s0 = p.res1 + g0 * p.ldr1 + 2 * t4;
                        const float* s1 = p.res1 + g1 * p.ldr1 + 2 * t4;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            {
                                const float2 a = ok0 ? __ldg(reinterpret_cast<const float2*>(s0 + 8 * j)) : make_float2(0.f, 0.f);
                                const float2 cc = ok1 ? __ldg(reinterpret_cast<const float2*>(s1 + 8 * j)) : make_float2(0.f, 0.f);
                                v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += cc.x; v[4 * j + 3] += cc.y;
                            }
                        }
                    }
                    if (p.fuse_u) {
                        // Fuse: stride-2 transposed-conv scatter of U (networks.py:199-206); tt1 = tt0 + 8 has tt0's parity
                        for (int tau = tt0 & 1; tau < p.fuse_k; tau += 2) {
                            const int j0 = (tt0 - tau) >> 1, j1 = (tt1 - tau) >> 1;
                            const bool v0 = ok0 && tt0 >= tau && j0 < p.fuse_n1, v1 = ok1 && tt1 >= tau && j1 < p.fuse_n1;
                            const float* u0 = p.fuse_u + ((size_t)b * p.fuse_n1 + (v0 ? j0 : 0)) * p.fuse_ld + tau * N + 2 * t4;
                            const float* u1 = p.fuse_u + ((size_t)b * p.fuse_n1 + (v1 ? j1 : 0)) * p.fuse_ld + tau * N + 2 * t4;
#pragma unroll
                            for (int j = 0; j < NJ; ++j) {
                                const float2 a = v0 ? __ldg(reinterpret_cast<const float2*>(u0 + 8 * j)) : make_float2(0.f, 0.f);
                                const float2 cc = v1 ? __ldg(reinterpret_cast<const float2*>(u1 + 8 * j)) : make_float2(0.f, 0.f);
                                v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += cc.x; v[4 * j + 3] += cc.y;
                            }
                        }
                    }
                    if (p.ln_g) {
                        float sa = 0.f, sb = 0.f, qa = 0.f, qb = 0.f;
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            sa += v[4 * j] + v[4 * j + 1]; sb += v[4 * j + 2] + v[4 * j + 3];
                            qa = fmaf(v[4 * j], v[4 * j], qa); qa = fmaf(v[4 * j + 1], v[4 * j + 1], qa);
                            qb = fmaf(v[4 * j + 2], v[4 * j + 2], qb); qb = fmaf(v[4 * j + 3], v[4 * j + 3], qb);
                        }
                        float r0, n0, r1, n1;
                        quad_stats(sa, qa, inv_n, r0, n0);
                        quad_stats(sb, qb, inv_n, r1, n1);
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            const float2 gg = *reinterpret_cast<const float2*>(s_lng + 8 * j + 2 * t4);
                            const float2 bb = *reinterpret_cast<const float2*>(s_lnb + 8 * j + 2 * t4);
                            v[4 * j] = fmaf(fmaf(v[4 * j], r0, n0), gg.x, bb.x);
                            v[4 * j + 1] = fmaf(fmaf(v[4 * j + 1], r0, n0), gg.y, bb.y);
                            v[4 * j + 2] = fmaf(fmaf(v[4 * j + 2], r1, n1), gg.x, bb.x);
                            v[4 * j + 3] = fmaf(fmaf(v[4 * j + 3], r1, n1), gg.y, bb.y);
                        }
                    }
                    if (p.act2 == ACT_RELU) {
#pragma unroll
                        for (int k = 0; k < 4 * NJ; ++k) v[k] = fmaxf(v[k], 0.f);
                    }
                }
                if (p.Y) {
                    const bool z0 = (p.row_mask && ok0 && p.row_mask[g0]) || (p.zero_from && tt0 >= p.zero_from[b]);
                    const bool z1 = (p.row_mask && ok1 && p.row_mask[g1]) || (p.zero_from && tt1 >= p.zero_from[b]);
                    float* y0 = p.Y + g0 * p.ldy + cbase + 2 * t4;
                    float* y1 = p.Y + g1 * p.ldy + cbase + 2 * t4;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        {
                            if (ok0) *reinterpret_cast<float2*>(y0 + 8 * j) = z0 ? make_float2(0.f, 0.f) : make_float2(v[4 * j], v[4 * j + 1]);
                            if (ok1) *reinterpret_cast<float2*>(y1 + 8 * j) = z1 ? make_float2(0.f, 0.f) : make_float2(v[4 * j + 2], v[4 * j + 3]);
                        }
                    }
                }
            }
        }
    }

    if (failed) atomicExch(up.err, 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

// Returns -1 when the layer is outside the tensor-core kernel's envelope (caller uses the SIMT path).
// `count` (1..3) problems of identical geometry share ONE launch.
int launch_umma_rowgemm_batch(const RowGemmParams* ps, const void* const* w_h16, int count, cudaStream_t s) {
    if (count < 1 || count > 3) return -1;
    const RowGemmParams& p = ps[0];
    for (int i = 0; i < count; ++i) {
        const RowGemmParams& o = ps[i];
        if (!w_h16[i] || o.mode != ROW_PLAIN || o.res2) return -1;
        if (o.B != p.B || o.n_in != p.n_in || o.n_out != p.n_out || o.K != p.K || o.Nout != p.Nout || o.taps != p.taps ||
            o.stride != p.stride || o.pad != p.pad) return -1;
        if (o.Nout > 128 && (o.ln_g || o.dot_out || o.res1 || o.fuse_u || o.act2 != ACT_NONE)) return -1;
        if (o.fuse_u && (o.fuse_ld % 2 || o.taps != 1 || o.stride != 1)) return -1;
        if (o.lda % 4 || (o.Y && o.ldy % 2) || (o.res1 && o.ldr1 % 2)) return -1;
        if (o.act2 != ACT_NONE && o.act2 != ACT_RELU) return -1;
        if (o.act1 == ACT_TANH) return -1;
    }
    if (p.K % 16 || p.K > KMAX || p.K < 16) return -1;
    if (!((p.stride == 1 && (p.taps == 1 || p.taps == 3)) || (p.stride == 2 && p.taps == 1))) return -1;
    if (p.Nout % 8 || p.Nout < 16 || p.Nout > NPAR) return -1;
    if (p.Nout > 128 && p.Nout % 128) return -1;
    if ((size_t)p.taps * p.Nout * p.K * 4 > W_MAX) return -1;
    if (p.stride == 1 && p.n_in != p.n_out) return -1;       // 'same' convs only (pad = taps / 2)
    int* err_flag = umma_err_flag();
    ES_CHECK(err_flag, "cannot allocate the device error flag");
    static PerDeviceSlot<int> n_sm_once; int& n_sm = n_sm_once.get();
    if (!n_sm) {
        int dev = 0;
        ES_CUDA(cudaGetDevice(&dev));
        ES_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    const int nj = p.Nout > 128 ? 16 : p.Nout / 8;
    if (nj != 4 && nj != 8 && nj != 12 && nj != 16) return -1;
    static PerDeviceSlot<bool> attr_once; bool& attr_set = attr_once.get();   // function attributes are per device
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(umma_rowgemm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        ES_CUDA(cudaFuncSetAttribute(umma_rowgemm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        ES_CUDA(cudaFuncSetAttribute(umma_rowgemm_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        ES_CUDA(cudaFuncSetAttribute(umma_rowgemm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    UrgParams up;
    for (int i = 0; i < 3; ++i) {
        up.g[i] = ps[i < count ? i : 0];
        up.w_h16[i] = w_h16[i < count ? i : 0];
    }
    up.nprob = count;
    up.err = err_flag;
    const int n_tiles = p.B * ((p.n_out + TM - 1) / TM);
    int per_prob = n_sm / count;
    if (per_prob > n_tiles) per_prob = n_tiles;
    const int grid = per_prob * count;
    switch (nj) {
        case 4: ES_CUDA(launch_pdl(umma_rowgemm_kernel<4>, grid, NTHR, SMEM_BYTES, s, up)); break;
        case 8: ES_CUDA(launch_pdl(umma_rowgemm_kernel<8>, grid, NTHR, SMEM_BYTES, s, up)); break;
        case 12: ES_CUDA(launch_pdl(umma_rowgemm_kernel<12>, grid, NTHR, SMEM_BYTES, s, up)); break;
        default: ES_CUDA(launch_pdl(umma_rowgemm_kernel<16>, grid, NTHR, SMEM_BYTES, s, up)); break;
    }
    ES_LAUNCH_OK();
    return 0;
}

int launch_umma_rowgemm(const RowGemmParams& p, const void* w_h16, cudaStream_t s) {
    return launch_umma_rowgemm_batch(&p, &w_h16, 1, s);
}

}  // namespace es

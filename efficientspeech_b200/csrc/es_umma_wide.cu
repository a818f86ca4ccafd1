// K-streamed tcgen05 row GEMM for the phoneme-side layers that do not fit the resident-weight
// kernel of es_umma_enc.cu (small / base variants: K up to 1024, N up to 3072, dense k=3 convs
// whose taps do not fit in shared memory).  Same pipeline as es_umma_dec256.cu:
//
//   x ring   (4 stages)  [128 + taps - 1 rows][32 ch] fp32 via 16-byte cp.async (zero fill outside
//                        the utterance), three chunks in flight
//   A ring   (3 stages)  the chunk split into fp16 hi/lo, UMMA canonical K-major no-swizzle with
//                        ROW PANELS (rows 16 bytes apart), so the operand of conv tap tau is the same
//                        buffer with the descriptor start address advanced by tau * 16 bytes
//   W ring   (3 stages)  one "unit" = the [NT][32] slice of one tap (hi and lo planes), one bulk copy;
//                        the image is packed in exactly the order the kernel consumes it:
//                        [N / NT][K / 32][taps][2][4][NT][8] halves (packing.canon_split_units)
//
// 6 x tcgen05.mma (M128 x NT x K16, hi*hi + hi*lo + lo*hi, two K steps) per unit into one of two
// TMEM accumulators.  Column tiles of NT = 128 / 256 outputs are independent CTAs (blockIdx.y); a
// LayerNorm / scalar-head epilogue needs the whole row in one tile (Nout == NT).
//
// Epilogue (8 warps x 16 rows, mma-fragment layout, 64 columns per step): bias, boundary-aware tap
// bias, ReLU / exact-erf GELU, scalar head, residual, then either the final store or -- with
// LayerNorm -- statistics + write-back into the accumulator and a second pass (normalise, ReLU,
// padding mask, store), exactly the op order of the fp32 SIMT kernel (es_rowgemm.cu).
#include <stdlib.h>

#include "es_common.cuh"
#include "es_kernels.cuh"
#include "es_umma.cuh"

namespace es {
namespace {

using namespace umma;

constexpr int TMW = 128;                  // rows per tile (UMMA M)
constexpr int KC = 32;                    // channels per K chunk
constexpr int MAXTAPS = 3;
constexpr int NTHR = 416;                 // 13 warps: 0..7 epilogue, 8..11 producer, 12 issue
constexpr int NPROD = 128;
constexpr int NXS = 4;                    // x ring depth
constexpr int NAS = 3;                    // A ring depth
constexpr int NWS = 3;                    // W ring depth
constexpr int NMAX = 256;
constexpr int XROWS_MAX = TMW + MAXTAPS - 1;                // 130
constexpr uint32_t X_STAGE = XROWS_MAX * KC * 4;            // 16640
constexpr uint32_t A_PANEL = 131 * 16;                      // 2096: 8 channels of all staged rows (48 mod 128: bank spread)
constexpr uint32_t A_PLANE = (KC / 8) * A_PANEL;            // 8384
constexpr uint32_t A_STAGE = 2 * A_PLANE;                   // 16768 (hi, lo)
constexpr uint32_t W_STAGE = 2 * (KC / 8) * NMAX * 16;      // 32768 (hi, lo) for NT = 256

constexpr uint32_t OFF_X = 0;
constexpr uint32_t OFF_A = OFF_X + NXS * X_STAGE;           // 66560
constexpr uint32_t OFF_W = OFF_A + NAS * A_STAGE;           // 116864
constexpr uint32_t OFF_PAR = OFF_W + NWS * W_STAGE;         // bias, tapb[3], dotw, lng, lnb (7 x 256 floats)
constexpr uint32_t OFF_BAR = OFF_PAR + 7 * NMAX * 4;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 192;
static_assert(OFF_A % 128 == 0 && OFF_W % 128 == 0, "operand alignment");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct WideParams {
    RowGemmParams g;
    const void* w_units;
    int* err;
};

template <int NT, int TAPS>
__global__ void __launch_bounds__(NTHR, 1)
umma_wide_kernel(const WideParams wp) {
    const RowGemmParams& p = wp.g;
    constexpr int PAD = TAPS / 2;
    constexpr int XROWS = TMW + TAPS - 1;
    constexpr int XITER = (XROWS + 15) / 16;
    constexpr int NH = NT / 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* par = reinterpret_cast<float*>(smem + OFF_PAR);
    float* s_bias = par;
    float* s_tapb = par + NMAX;                              // [3][256]
    float* s_dotw = par + 4 * NMAX;
    float* s_lng = par + 5 * NMAX;
    float* s_lnb = par + 6 * NMAX;
    const uint32_t bar0 = smem_u32(smem + OFF_BAR);
    const uint32_t bar_wfull = bar0;            // [3] W unit landed
    const uint32_t bar_wfree = bar0 + 24;       // [3] W unit consumed by its MMAs
    const uint32_t bar_aready = bar0 + 48;      // [3] A chunk written by the 4 producer warps
    const uint32_t bar_afree = bar0 + 72;       // [3] A chunk consumed by all its taps
    const uint32_t bar_accfull = bar0 + 96;     // [2]
    const uint32_t bar_accfree = bar0 + 112;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 128);

    const int K = p.K;
    const int nchunks = K / KC;
    const int col0 = blockIdx.y * NT;                        // first output column of this CTA
    const int tiles_per_utt = (p.n_out + TMW - 1) / TMW;
    const int n_tiles = p.B * tiles_per_utt;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_chunks = my_tiles * nchunks;
    const int total_units = total_chunks * TAPS;
    constexpr uint32_t w_plane = (uint32_t)(KC / 8) * NT * 16u;
    constexpr uint32_t w_unit_bytes = 2 * w_plane;
    const int units_per_tile = nchunks * TAPS;
    const uint8_t* w_base_g = reinterpret_cast<const uint8_t*>(wp.w_units) + (size_t)blockIdx.y * units_per_tile * w_unit_bytes;

    // ---- one-time setup ---------------------------------------------------------------------
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
    if (tid == 0) {
        for (int k = 0; k < 3; ++k) {
            mbar_init(bar_wfull + 8 * k, 1);
            mbar_init(bar_wfree + 8 * k, 1);
            mbar_init(bar_aready + 8 * k, 4);
            mbar_init(bar_afree + 8 * k, 1);
        }
        for (int k = 0; k < 2; ++k) {
            mbar_init(bar_accfull + 8 * k, 1);
            mbar_init(bar_accfree + 8 * k, 8);
        }
        fence_mbar_init();
    }
    for (int i = tid; i < NMAX; i += NTHR) {
        const bool in = i < NT;
        s_bias[i] = (p.bias && in) ? __ldg(p.bias + col0 + i) : 0.f;
        for (int t = 0; t < MAXTAPS; ++t)
            s_tapb[t * NMAX + i] = (p.tap_bias && t < TAPS && in) ? __ldg(p.tap_bias + t * p.ldw + col0 + i) : 0.f;
        s_dotw[i] = (p.dot_w && in) ? __ldg(p.dot_w + i) : 0.f;
        s_lng[i] = (p.ln_g && in) ? __ldg(p.ln_g + i) : 0.f;
        s_lnb[i] = (p.ln_g && in) ? __ldg(p.ln_b + i) : 0.f;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    bool failed = false;
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 12) {
        // =========================================================================== issue warp
        const bool elected = elect_one();
        constexpr uint32_t idesc = make_idesc_f16(TMW, NT);
        constexpr uint32_t lbo_b = (uint32_t)NT * 16u;
        auto load_w = [&](int st, int u) {       // unit u of the tile's weight stream -> stage st
            mbar_arrive_expect_tx(bar_wfull + 8 * st, w_unit_bytes);
            bulk_g2s(smem_u32(smem + OFF_W) + (uint32_t)st * W_STAGE, w_base_g + (size_t)u * w_unit_bytes, w_unit_bytes,
                     bar_wfull + 8 * st);
        };
        if (elected) {
            if (total_units > 0) load_w(0, 0);
            if (total_units > 1) load_w(1, 1 % units_per_tile);
        }
        __syncwarp();
        int i = 0, c = 0, t = 0;                 // tile, chunk, tap of unit g
        int ws = 0, wuse = 0;                    // W ring stage / pass of unit g
        int ws2 = 2 % NWS, u2 = 2 % units_per_tile;   // stage / unit index of unit g + 2
        int as = 0, ause = 0;                    // A ring stage / pass of the current chunk
        for (int g = 0; g < total_units; ++g) {
            const int acc = i & 1;
            if (g + 2 < total_units) {           // prefetch two units ahead; that stage was last used by unit g-1
                if (g >= 1 && !mbar_wait(bar_wfree + 8 * ws2, ((g - 1) / NWS) & 1)) failed = true;
                if (elected) load_w(ws2, u2);
            }
            if (!mbar_wait(bar_wfull + 8 * ws, wuse & 1)) failed = true;
            if (t == 0) {
                if (!mbar_wait(bar_aready + 8 * as, ause & 1)) failed = true;
                if (c == 0 && i >= 2 && !mbar_wait(bar_accfree + 8 * acc, ((i >> 1) - 1) & 1)) failed = true;
            }
            tc_fence_after_sync();
            const uint32_t a_base = smem_u32(smem + OFF_A) + (uint32_t)as * A_STAGE + (uint32_t)t * 16u;   // tap: row shift
            const uint32_t w_base = smem_u32(smem + OFF_W) + (uint32_t)ws * W_STAGE;
            const uint32_t d = tmem + (uint32_t)(acc * NMAX);
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks) {
                const uint64_t dah = make_smem_desc(a_base + (uint32_t)(2 * ks) * A_PANEL, A_PANEL, 128u);
                const uint64_t dal = make_smem_desc(a_base + A_PLANE + (uint32_t)(2 * ks) * A_PANEL, A_PANEL, 128u);
                const uint64_t dbh = make_smem_desc(w_base + (uint32_t)(2 * ks) * lbo_b, lbo_b, 128u);
                const uint64_t dbl = make_smem_desc(w_base + w_plane + (uint32_t)(2 * ks) * lbo_b, lbo_b, 128u);
                if (elected) {
                    mma_f16_ss(d, dah, dbh, idesc, (c > 0 || t > 0 || ks > 0) ? 1u : 0u);
                    mma_f16_ss(d, dah, dbl, idesc, 1u);
                    mma_f16_ss(d, dal, dbh, idesc, 1u);
                }
            }
            if (elected) {
                mma_commit(bar_wfree + 8 * ws);
                if (t == TAPS - 1) {
                    mma_commit(bar_afree + 8 * as);
                    if (c == nchunks - 1) mma_commit(bar_accfull + 8 * acc);
                }
            }
            __syncwarp();
            if (++ws == NWS) { ws = 0; ++wuse; }
            if (++ws2 == NWS) ws2 = 0;
            if (++u2 == units_per_tile) u2 = 0;
            if (++t == TAPS) {
                t = 0;
                if (++as == NAS) { as = 0; ++ause; }
                if (++c == nchunks) { c = 0; ++i; }
            }
        }
    } else if (warp >= 8) {
        // =========================================================================== producers
        const int ptid = tid - 256;
        const int q = ptid & 7;                               // 16-byte column piece (4 channels) of the chunk
        const int xrow = ptid >> 3;                           // rows xrow + 16*it
        const uint32_t x_smem = smem_u32(smem + OFF_X);

        int ld_i = 0, ld_c = 0, ld_s = 0;
        const float* ld_base = p.A;
        uint32_t ld_mask = 0;
        auto tile_setup = [&](int i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TMW;
            const int tf = t0 - PAD + xrow;
            ld_base = p.A + ((long long)b * p.n_in + tf) * p.lda + q * 4;
            ld_mask = 0;
#pragma unroll
            for (int it = 0; it < XITER; ++it) {
                const int t = tf + 16 * it;
                if (t >= 0 && t < p.n_in && xrow + 16 * it < XROWS) ld_mask |= 1u << it;
            }
        };
        auto load_next = [&]() {
            const uint32_t dst = x_smem + (uint32_t)ld_s * X_STAGE + (uint32_t)ptid * 16u;
#pragma unroll
            for (int it = 0; it < XITER; ++it) {
                if (XROWS % 16 == 0 || it < XITER - 1 || xrow + 16 * it < XROWS) {
                    const bool ok = (ld_mask >> it) & 1u;
                    const float* src = ok ? ld_base + (size_t)(16 * it) * p.lda + ld_c * KC : p.A;
                    cp_async16(dst + (uint32_t)it * (16u * KC * 4u), src, ok ? 16u : 0u);
                }
            }
            if (++ld_s == NXS) ld_s = 0;
            if (++ld_c == nchunks) {
                ld_c = 0;
                if (++ld_i < my_tiles) tile_setup(ld_i);
            }
        };
        if (my_tiles > 0) tile_setup(0);
        for (int g = 0; g < 3; ++g) {
            if (g < total_chunks) load_next();
            cp_async_commit();
        }
        int sx = 0, sa = 0, use = 0;
        for (int g = 0; g < total_chunks; ++g) {
            cp_async_wait<2>();
            named_bar_sync(2, NPROD);
            if (g + 3 < total_chunks) load_next();
            cp_async_commit();

            const uint8_t* Xc = smem + OFF_X + (uint32_t)sx * X_STAGE + (uint32_t)ptid * 16u;
            uint2 hi[XITER], lo[XITER];
#pragma unroll
            for (int it = 0; it < XITER; ++it) {
                if (XROWS % 16 == 0 || it < XITER - 1 || xrow + 16 * it < XROWS)
                    split4(*reinterpret_cast<const float4*>(Xc + it * (16 * KC * 4)), hi[it], lo[it]);
            }
            if (g >= NAS && !mbar_wait(bar_afree + 8 * sa, (use - 1) & 1)) failed = true;
            uint8_t* a_hi = smem + OFF_A + (uint32_t)sa * A_STAGE
                            + (uint32_t)(q >> 1) * A_PANEL + (uint32_t)xrow * 16u + (uint32_t)(q & 1) * 8u;
#pragma unroll
            for (int it = 0; it < XITER; ++it) {
                if (XROWS % 16 == 0 || it < XITER - 1 || xrow + 16 * it < XROWS) {
                    *reinterpret_cast<uint2*>(a_hi + it * 256) = hi[it];
                    *reinterpret_cast<uint2*>(a_hi + A_PLANE + it * 256) = lo[it];
                }
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_aready + 8 * sa);
            if (++sx == NXS) sx = 0;
            if (++sa == NAS) { sa = 0; ++use; }
        }
        cp_async_wait<0>();
    } else {
        // =========================================================================== epilogue
        const int qd = warp & 3, half = warp >> 2;
        const int rbase = qd * 32 + half * 16;
        const int t4 = lane & 3, tr = lane >> 2;
        const float inv_n = 1.f / (float)p.Nout;
        const uint32_t lane_addr = (uint32_t)rbase << 16;
        const bool has_ln = p.ln_g != nullptr;

        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TMW;
            const int rows_valid = min(TMW, p.n_out - t0);
            const int acc = i & 1;
            const int row0 = rbase + tr, row1 = row0 + 8;
            const int tt0 = t0 + row0, tt1 = t0 + row1;
            const size_t g0 = (size_t)b * p.n_out + tt0, g1 = g0 + 8;
            const bool ok0 = row0 < rows_valid, ok1 = row1 < rows_valid;
            const uint32_t tacc = tmem + lane_addr + (uint32_t)(acc * NMAX);
            const bool z0 = (p.row_mask && ok0 && p.row_mask[g0]) || (p.zero_from && tt0 >= p.zero_from[b]);
            const bool z1 = (p.row_mask && ok1 && p.row_mask[g1]) || (p.zero_from && tt1 >= p.zero_from[b]);
            // tap-bias masks: tap t of output row tt reads input row tt + t - PAD
            float m0[TAPS], m1[TAPS];
#pragma unroll
            for (int t = 0; t < TAPS; ++t) {
                const int ti0 = tt0 + t - PAD, ti1 = tt1 + t - PAD;
                m0[t] = (ti0 >= 0 && ti0 < p.n_in) ? 1.f : 0.f;
                m1[t] = (ti1 >= 0 && ti1 < p.n_in) ? 1.f : 0.f;
            }
            if (!mbar_wait(bar_accfull + 8 * acc, (i >> 1) & 1)) failed = true;
            tc_fence_after_sync();

            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll 1
            for (int h = 0; h < NH; ++h) {
                uint32_t r[32];
                tmem_ld_16x256b_x8(tacc + (uint32_t)(h * 64), r);
                const int cb = h * 64 + 2 * t4;                 // this thread's first column of the step (tile-local)
                float2 k0[8], k1[8];
                if (p.res1) {
                    const float* sp0 = p.res1 + g0 * p.ldr1 + col0 + cb;
                    const float* sp1 = p.res1 + g1 * p.ldr1 + col0 + cb;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        k0[j] = ok0 ? __ldg(reinterpret_cast<const float2*>(sp0 + 8 * j)) : make_float2(0.f, 0.f);
                        k1[j] = ok1 ? __ldg(reinterpret_cast<const float2*>(sp1 + 8 * j)) : make_float2(0.f, 0.f);
                    }
                }
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 bb = *reinterpret_cast<const float2*>(s_bias + cb + 8 * j);
                    v[4 * j] = __uint_as_float(r[4 * j]) + bb.x; v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + bb.y;
                    v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + bb.x; v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + bb.y;
                }
                if (p.tap_bias) {
#pragma unroll
                    for (int t = 0; t < TAPS; ++t) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 tb = *reinterpret_cast<const float2*>(s_tapb + t * NMAX + cb + 8 * j);
                            v[4 * j] = fmaf(m0[t], tb.x, v[4 * j]); v[4 * j + 1] = fmaf(m0[t], tb.y, v[4 * j + 1]);
                            v[4 * j + 2] = fmaf(m1[t], tb.x, v[4 * j + 2]); v[4 * j + 3] = fmaf(m1[t], tb.y, v[4 * j + 3]);
                        }
                    }
                }
                if (p.act1 == ACT_RELU) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] = fmaxf(v[k], 0.f);
                } else if (p.act1 == ACT_GELU) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] = gelu_erf_f(v[k]);
                }
                if (p.dot_out) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float2 w = *reinterpret_cast<const float2*>(s_dotw + cb + 8 * j);
                        d0 = fmaf(v[4 * j], w.x, d0); d0 = fmaf(v[4 * j + 1], w.y, d0);
                        d1 = fmaf(v[4 * j + 2], w.x, d1); d1 = fmaf(v[4 * j + 3], w.y, d1);
                    }
                }
                if (p.res1) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v[4 * j] += k0[j].x; v[4 * j + 1] += k0[j].y; v[4 * j + 2] += k1[j].x; v[4 * j + 3] += k1[j].y;
                    }
                }
                if (p.fuse_u) {
                    for (int tau = tt0 & 1; tau < p.fuse_k; tau += 2) {      // tt1 = tt0 + 8 has the same parity
                        const int j0 = (tt0 - tau) >> 1, j1 = (tt1 - tau) >> 1;
                        const bool v0 = ok0 && tt0 >= tau && j0 < p.fuse_n1, v1 = ok1 && tt1 >= tau && j1 < p.fuse_n1;
                        const float* u0 = p.fuse_u + ((size_t)b * p.fuse_n1 + j0) * p.fuse_ld + tau * p.Nout + col0 + cb;
                        const float* u1 = p.fuse_u + ((size_t)b * p.fuse_n1 + j1) * p.fuse_ld + tau * p.Nout + col0 + cb;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 a = v0 ? __ldg(reinterpret_cast<const float2*>(u0 + 8 * j)) : make_float2(0.f, 0.f);
                            const float2 c2 = v1 ? __ldg(reinterpret_cast<const float2*>(u1 + 8 * j)) : make_float2(0.f, 0.f);
                            v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += c2.x; v[4 * j + 3] += c2.y;
                        }
                    }
                }
                if (has_ln) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        s0 += v[4 * j] + v[4 * j + 1]; q0 = fmaf(v[4 * j], v[4 * j], q0); q0 = fmaf(v[4 * j + 1], v[4 * j + 1], q0);
                        s1 += v[4 * j + 2] + v[4 * j + 3]; q1 = fmaf(v[4 * j + 2], v[4 * j + 2], q1); q1 = fmaf(v[4 * j + 3], v[4 * j + 3], q1);
                    }
#pragma unroll
                    for (int k = 0; k < 32; ++k) r[k] = __float_as_uint(v[k]);
                    tmem_st_16x256b_x8(tacc + (uint32_t)(h * 64), r);
                } else if (p.Y) {
                    if (p.act2 == ACT_RELU) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) v[k] = fmaxf(v[k], 0.f);
                    }
                    float* y0 = p.Y + g0 * p.ldy + col0 + cb;
                    float* y1 = p.Y + g1 * p.ldy + col0 + cb;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (ok0) *reinterpret_cast<float2*>(y0 + 8 * j) = z0 ? make_float2(0.f, 0.f) : make_float2(v[4 * j], v[4 * j + 1]);
                        if (ok1) *reinterpret_cast<float2*>(y1 + 8 * j) = z1 ? make_float2(0.f, 0.f) : make_float2(v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            }
            if (p.dot_out) {
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
            if (has_ln) {
                tmem_st_wait();
                float ra, na, rb, nb;
                quad_stats(s0, q0, inv_n, ra, na);
                quad_stats(s1, q1, inv_n, rb, nb);
                if (p.Y) {
#pragma unroll 1
                    for (int h = 0; h < NH; ++h) {
                        uint32_t r[32];
                        tmem_ld_16x256b_x8(tacc + (uint32_t)(h * 64), r);
                        tmem_ld_wait();
                        const int cb = h * 64 + 2 * t4;
                        float* y0 = p.Y + g0 * p.ldy + col0 + cb;
                        float* y1 = p.Y + g1 * p.ldy + col0 + cb;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float2 gg = *reinterpret_cast<const float2*>(s_lng + cb + 8 * j);
                            const float2 bb = *reinterpret_cast<const float2*>(s_lnb + cb + 8 * j);
                            float a = fmaf(fmaf(__uint_as_float(r[4 * j]), ra, na), gg.x, bb.x);
                            float bq = fmaf(fmaf(__uint_as_float(r[4 * j + 1]), ra, na), gg.y, bb.y);
                            float cq = fmaf(fmaf(__uint_as_float(r[4 * j + 2]), rb, nb), gg.x, bb.x);
                            float dq = fmaf(fmaf(__uint_as_float(r[4 * j + 3]), rb, nb), gg.y, bb.y);
                            if (p.act2 == ACT_RELU) { a = fmaxf(a, 0.f); bq = fmaxf(bq, 0.f); cq = fmaxf(cq, 0.f); dq = fmaxf(dq, 0.f); }
                            if (ok0) *reinterpret_cast<float2*>(y0 + 8 * j) = z0 ? make_float2(0.f, 0.f) : make_float2(a, bq);
                            if (ok1) *reinterpret_cast<float2*>(y1 + 8 * j) = z1 ? make_float2(0.f, 0.f) : make_float2(cq, dq);
                        }
                    }
                }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accfree + 8 * acc);
        }
    }

    if (failed) atomicExch(wp.err, 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int NT, int TAPS>
int launch_wide(const WideParams& wp, dim3 grid, cudaStream_t s) {
    static PerDeviceSlot<bool> attr_once; bool& attr_set = attr_once.get();   // function attributes are per device
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(umma_wide_kernel<NT, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    ES_CUDA(launch_pdl(umma_wide_kernel<NT, TAPS>, grid, NTHR, SMEM_BYTES, s, wp));
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace

// Which split-fp16 weight image a dense phoneme-side layer uses (the host packs accordingly):
// 0 none (fp32 SIMT only), 1 resident taps image (es_umma_enc.cu), 2 / 3 streamed units with
// 128 / 256 output columns per tile (this file).
// output columns per CTA of the streamed kernel for this geometry (0: not applicable)
static int wide_tile(int K, int Nout, int taps, int stride) {
    if (stride != 1 || (taps != 1 && taps != 3) || K % KC || K < KC) return 0;
    return Nout % 256 == 0 ? 256 : Nout % 128 == 0 ? 128 : 0;
}

int dense_layout(int K, int Nout, int taps, int stride) {
    const bool geom1 = (stride == 1 && (taps == 1 || taps == 3)) || (stride == 2 && taps == 1);
    if (geom1 && K % 16 == 0 && K >= 16 && K <= 128 && Nout % 8 == 0 && Nout >= 16 && Nout <= 384 &&
        (Nout <= 128 || Nout % 128 == 0) && (size_t)taps * Nout * K * 4 <= 100 * 1024) {
        const int nj = Nout > 128 ? 16 : Nout / 8;
        if (nj == 4 || nj == 8 || nj == 12 || nj == 16) return 1;
    }
    const int nt = wide_tile(K, Nout, taps, stride);
    return nt == 256 ? 3 : nt == 128 ? 2 : 0;
}

// Returns -1 when the layer's epilogue is outside this kernel's envelope (caller uses the SIMT path).
// `nt_force` (128): the image was packed with that tile width regardless of dense_layout (Fuse).
int launch_umma_wide(const RowGemmParams& p_in, const void* w_units, cudaStream_t s, int nt_force) {
    int NT = wide_tile(p_in.K, p_in.Nout, p_in.taps, p_in.stride);
    if (nt_force) {
        if (NT == 0 || p_in.Nout % nt_force) return -1;
        NT = nt_force;
    } else if (dense_layout(p_in.K, p_in.Nout, p_in.taps, p_in.stride) < 2) {
        return -1;
    }
    if (NT == 0 || !w_units) return -1;
    RowGemmParams p = p_in;
    if (p.mode != ROW_PLAIN || p.res2 || p.act1 == ACT_TANH) return -1;
    if (p.act2 != ACT_NONE && p.act2 != ACT_RELU) return -1;
    if ((p.ln_g || p.dot_out) && p.Nout != NT) return -1;      // whole row in one column tile
    if (p.lda % 4 || (p.Y && p.ldy % 2) || (p.res1 && p.ldr1 % 2)) return -1;
    if (p.n_in != p.n_out || p.pad != p.taps / 2) return -1;   // 'same' convs only
    if (p.taps == 1 && !p.zero_from && !p.fuse_u) {            // row-local: utterance boundaries do not matter
        p.n_in = p.n_out = p.B * p.n_out;
        p.B = 1;
    }
    int* err_flag = umma_err_flag();
    ES_CHECK(err_flag, "cannot allocate the device error flag");
    static PerDeviceSlot<int> n_sm_once; int& n_sm = n_sm_once.get();
    if (!n_sm) {
        int dev = 0;
        ES_CUDA(cudaGetDevice(&dev));
        ES_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    WideParams wp;
    wp.g = p; wp.w_units = w_units; wp.err = err_flag;
    const int n_tiles = p.B * ((p.n_out + TMW - 1) / TMW);
    const int ny = p.Nout / NT;
    int gx = (n_sm + ny - 1) / ny;
    if (gx > n_tiles) gx = n_tiles;
    const dim3 grid(gx, ny);
    if (NT == 256) return p.taps == 3 ? launch_wide<256, 3>(wp, grid, s) : launch_wide<256, 1>(wp, grid, s);
    return p.taps == 3 ? launch_wide<128, 3>(wp, grid, s) : launch_wide<128, 1>(wp, grid, s);
}

}  // namespace es

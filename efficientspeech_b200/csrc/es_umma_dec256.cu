// Decoder-side tensor-core kernel for the WIDE decoders (small / base: dx2 = 256 channels).
//
// The split-fp16 image of a 256x256 pointwise weight is 256 KB -- it cannot stay resident next to
// the activations as in es_umma_dec.cu.  This kernel therefore streams EVERYTHING along K:
// a 128-frame tile is processed in chunks of 32 input channels, and three rings run in lock step
//
//   x ring   (4 stages)  [132 rows][32 ch] fp32, filled by the producers with 16-byte cp.async
//                        (zero-fill outside the utterance / for padded frames), 3 chunks in flight.
//                        Every producer thread owns one 16-byte column piece and rows r, r+16, ...:
//                        the per-chunk address work is one pointer plus compile-time offsets.
//   A ring   (3 stages)  the producers' depthwise conv (k=5, sliding window in registers) of the
//                        chunk, split fp16 hi/lo, UMMA canonical K-major no-swizzle panels.  The
//                        panel stride (descriptor LBO) is padded by 16 bytes so that the four
//                        panels a warp writes at once fall into different banks.
//   W ring   (3 stages)  the matching [N][32] slice of the split weights (pre-chunked at pack
//                        time), one bulk async copy (TMA 1-D) per chunk, from L2
//
// and the issue warp launches 6 x tcgen05.mma (M128 x N x K16: hi*hi + hi*lo + lo*hi for two K
// steps) per chunk into one of two 256-column TMEM accumulators; one tcgen05.commit per chunk
// releases the A/W stage, the last chunk's commit hands the accumulator to the epilogue.
//
// Epilogue (8 warps, 16 rows each, mma-fragment layout, 4 threads per row): a 256-channel row is
// 2 x 64 values per thread -- too many to hold -- so the row statistics are taken in passes over
// the accumulator, 64 columns (32 registers) at a time with the next TMEM load in flight, and
// the accumulator doubles as scratch: tanh(acc + bias) is written back with tcgen05.st,
// normalised (+ skip, second statistics, written back again on block-end layers) and finally
// stored with 8-byte accesses (8 rows x 32 contiguous bytes per warp instruction).
//
// Same modes as es_umma_dec.cu: DWCONV (decoder layer), PLAIN (per-phoneme projection with K = 4d up to 512;
// mel head, N = 80).
#include <stdlib.h>

#include "es_common.cuh"
#include "es_kernels.cuh"
#include "es_umma.cuh"

namespace es {
namespace {

using namespace umma;

constexpr int TM2 = 128;                  // frames per tile (UMMA M)
constexpr int KC = 32;                    // channels per K chunk
constexpr int DWK = 5;
constexpr int NTHR = 416;                 // 13 warps: 0..7 epilogue, 8..11 producer, 12 issue
constexpr int NPROD = 128;
constexpr int NXS = 4;                    // x ring depth
constexpr int NAS = 3;                    // A / W ring depth
constexpr int NMAX = 256;
constexpr uint32_t X_STAGE = (TM2 + DWK - 1) * KC * 4;      // 16896
constexpr uint32_t A_PANEL = TM2 * 16 + 16;                 // 2064: 8 channels of all 128 rows (+16: bank spread)
constexpr uint32_t A_PLANE = (KC / 8) * A_PANEL;            // 8256
constexpr uint32_t A_STAGE = 2 * A_PLANE;                   // 16512 (hi, lo)
constexpr uint32_t W_STAGE = 2 * (KC / 8) * NMAX * 16;      // 32768 (hi, lo)

constexpr uint32_t OFF_X = 0;
constexpr uint32_t OFF_A = OFF_X + NXS * X_STAGE;           // 67584
constexpr uint32_t OFF_W = OFF_A + NAS * A_STAGE;           // 117120
constexpr uint32_t OFF_PAR = OFF_W + NAS * W_STAGE;         // bias, ln g/b, ln2 g/b (5 x 256 floats)
constexpr uint32_t OFF_DW = OFF_PAR + 5 * NMAX * 4;         // depthwise taps + bias (6 x 256 floats)
constexpr uint32_t OFF_SRC = OFF_DW + 6 * NMAX * 4;         // gather sources, 2 x 128 ints
constexpr uint32_t OFF_CUM = OFF_SRC + 2 * TM2 * 4;         // duration prefix sums of the utterance (<= 1024)
constexpr uint32_t OFF_BAR = OFF_CUM + 1024 * 4;
constexpr int TL_CACHE = 64;                                // ragged schedule: this CTA's first 64 tile coordinates
constexpr uint32_t OFF_TL = OFF_BAR + 128;
constexpr uint32_t SMEM_BYTES = OFF_TL + TL_CACHE * 8;
static_assert(OFF_W % 128 == 0 && OFF_A % 128 == 0, "operand alignment");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

enum { MODE_DWCONV = 0, MODE_PLAIN = 2 };

struct Dec256Params {
    int B, T, K;                 // K input channels (multiple of 32, <= 512)
    const float* X;              // [B,T,K]
    const float* dw_w;           // [5][K]
    const float* dw_b;           // [K]
    const void* w_chunks;        // [K/32][2 (hi,lo)][4][N][8] halves
    const float* bias;
    int act_tanh;
    const float* ln_g; const float* ln_b;
    const float* res2; const float* ln2_g; const float* ln2_b;
    const int* zero_from;
    float* Y;
    int* err;
    const int2* tile_list;       // ragged scheduling (es_gather.cu): (b, t0) of the tiles that can reach a valid frame, or null
    const int* tile_count;
    const int* src;              // GX / GS kernels: frame -> table row map [B*T] (es_gather.cu); X / res2 is the projection table
};

constexpr float kTanhScale2 = 2.8853900817779268f;
__device__ __forceinline__ float tanh_scaled2(float arg) {     // see es_umma_dec.cu
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(arg));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return fmaf(-2.f, r, 1.f);
}

// final store of one 64-column step of the fragment (2 rows x 16 columns per thread); NJ = valid
// 8-column groups.  ZERO: rows at or beyond the utterance's mel length store zeros.
template <int NJ, bool ZERO>
__device__ __forceinline__ void store_frag(float* y0, float* y1, bool ok0, bool ok1, bool z0, bool z1,
                                           const uint32_t (&r)[32]) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        float2 a = make_float2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]));
        float2 c = make_float2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
        if (ZERO) {
            if (z0) a = make_float2(0.f, 0.f);
            if (z1) c = make_float2(0.f, 0.f);
        }
        if (ok0) *reinterpret_cast<float2*>(y0 + 8 * j) = a;
        if (ok1) *reinterpret_cast<float2*>(y1 + 8 * j) = c;
    }
}

// y = LN(v) on one 64-column step: v*r + nm, then the affine pair at parc + 8j
__device__ __forceinline__ void norm_frag(uint32_t (&r)[32], const float* g, const float* be,
                                          float ra, float na, float rb, float nb) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float2 gg = *reinterpret_cast<const float2*>(g + 8 * j);
        const float2 bb = *reinterpret_cast<const float2*>(be + 8 * j);
        r[4 * j] = __float_as_uint(fmaf(fmaf(__uint_as_float(r[4 * j]), ra, na), gg.x, bb.x));
        r[4 * j + 1] = __float_as_uint(fmaf(fmaf(__uint_as_float(r[4 * j + 1]), ra, na), gg.y, bb.y));
        r[4 * j + 2] = __float_as_uint(fmaf(fmaf(__uint_as_float(r[4 * j + 2]), rb, nb), gg.x, bb.x));
        r[4 * j + 3] = __float_as_uint(fmaf(fmaf(__uint_as_float(r[4 * j + 3]), rb, nb), gg.y, bb.y));
    }
}

// GX: the input rows, GS: the skip rows of a block-end layer are rows p.src[b*T + t] of a table instead of rows
// b*T + t of a [B,T,256] tensor -- the length regulator fused into the first decoder block,
// separate instantiations because the dense kernel sits at its register limit.
template <int MODE, int N, bool GX = false, bool GS = false>
__global__ void __launch_bounds__(NTHR, 1)
umma_dec256_kernel(const Dec256Params p) {
    constexpr int HALO = (MODE == MODE_DWCONV) ? DWK / 2 : 0;
    constexpr int XROWS = TM2 + 2 * HALO;
    constexpr int XITER = (XROWS + 15) / 16;                   // rows r, r+16, ... per producer thread
    constexpr int NH = (N + 63) / 64;                          // 64-column epilogue steps
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* par = reinterpret_cast<float*>(smem + OFF_PAR);
    float* dws = reinterpret_cast<float*>(smem + OFF_DW);
    const uint32_t bar0 = smem_u32(smem + OFF_BAR);
    const uint32_t bar_wfull = bar0;            // [3] W chunk landed
    const uint32_t bar_cfree = bar0 + 24;       // [3] chunk stage (A and W) consumed by its MMAs
    const uint32_t bar_aready = bar0 + 48;      // [3] A chunk written by the 4 producer warps
    const uint32_t bar_accfull = bar0 + 72;     // [2]
    const uint32_t bar_accfree = bar0 + 88;     // [2] accumulator drained by the 8 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 112);

    const int K = (MODE == MODE_DWCONV) ? 256 : p.K;
    const int nchunks = K / KC;
    const int tiles_per_utt = (p.T + TM2 - 1) / TM2;
    // tile index -> (utterance, first frame): dense order, or the compacted list of the ragged schedule (es_gather.cu)
    const int2* s_tiles = reinterpret_cast<const int2*>(smem + OFF_TL);
    bool use_list = false;        // set after the dependency wait: a list that holds EVERY tile is the dense order itself
    auto tile_bt = [&](int tile, int& b, int& t0) {
        if (use_list) {
            const int k = (tile - (int)blockIdx.x) / (int)gridDim.x;      // staged in shared memory after the dependency wait
            const int2 v = k < TL_CACHE ? s_tiles[k] : __ldg(p.tile_list + tile);
            b = v.x; t0 = v.y;
        } else {
            b = tile / tiles_per_utt; t0 = (tile - b * tiles_per_utt) * TM2;
        }
    };
    constexpr uint32_t w_plane = (uint32_t)(KC / 8) * N * 16u;   // bytes of one fp16 plane of one chunk
    constexpr uint32_t w_chunk_bytes = 2 * w_plane;

    // ---- one-time setup ---------------------------------------------------------------------
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);      // two 256-column fp32 accumulators
    if (tid == 0) {
        for (int k = 0; k < NAS; ++k) {
            mbar_init(bar_wfull + 8 * k, 1);
            mbar_init(bar_cfree + 8 * k, 1);
            mbar_init(bar_aready + 8 * k, 4);
        }
        for (int k = 0; k < 2; ++k) {
            mbar_init(bar_accfull + 8 * k, 1);
            mbar_init(bar_accfree + 8 * k, 8);
        }
        fence_mbar_init();
    }
    for (int i = tid; i < NMAX; i += NTHR) {
        par[i] = (i < N) ? __ldg(p.bias + i) * (p.act_tanh ? kTanhScale2 : 1.f) : 0.f;
        par[NMAX + i] = (p.ln_g && i < N) ? __ldg(p.ln_g + i) : 0.f;
        par[2 * NMAX + i] = (p.ln_g && i < N) ? __ldg(p.ln_b + i) : 0.f;
        par[3 * NMAX + i] = (p.ln2_g && i < N) ? __ldg(p.ln2_g + i) : 0.f;
        par[4 * NMAX + i] = (p.ln2_g && i < N) ? __ldg(p.ln2_b + i) : 0.f;
    }
    if (MODE == MODE_DWCONV) {
        for (int i = tid; i < (DWK + 1) * NMAX; i += NTHR) {
            const int t = i / NMAX, c = i - t * NMAX;
            dws[i] = t < DWK ? __ldg(p.dw_w + t * K + c) : __ldg(p.dw_b + c);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    bool failed = false;
    pdl_launch_dependents();
    pdl_wait();
    const int n_tiles = p.tile_count ? *reinterpret_cast<const volatile int*>(p.tile_count) : p.B * tiles_per_utt;
    if (p.tile_count) {
        use_list = n_tiles != p.B * tiles_per_utt;           // CTA-uniform
        if (use_list) {
            if (tid < TL_CACHE) {
                const int tile = blockIdx.x + tid * gridDim.x;
                if (tile < n_tiles) reinterpret_cast<int2*>(smem + OFF_TL)[tid] = __ldg(p.tile_list + tile);
            }
            __syncthreads();
        }
    }
    const int my_tiles = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int total_chunks = my_tiles * nchunks;

    if (warp == 12) {
        // =========================================================================== issue warp
        const bool elected = elect_one();
        constexpr uint32_t idesc = make_idesc_f16(TM2, N);
        constexpr uint32_t lbo_b = (uint32_t)N * 16u;
        auto load_w = [&](int st, int c) {       // chunk c of the weight -> stage st
            mbar_arrive_expect_tx(bar_wfull + 8 * st, w_chunk_bytes);
            bulk_g2s(smem_u32(smem + OFF_W) + (uint32_t)st * W_STAGE,
                     reinterpret_cast<const uint8_t*>(p.w_chunks) + (size_t)c * w_chunk_bytes, w_chunk_bytes,
                     bar_wfull + 8 * st);
        };
        if (elected) {
            if (total_chunks > 0) load_w(0, 0);
            if (total_chunks > 1) load_w(1, 1 % nchunks);
        }
        __syncwarp();
        int i = 0, c = 0, st = 0, use = 0;       // tile, chunk, ring stage, ring pass of chunk g
        int st2 = 2 % NAS, c2 = 2 % nchunks;     // stage / weight chunk of chunk g + 2
        for (int g = 0; g < total_chunks; ++g) {
            const int acc = i & 1;
            // prefetch the weights two chunks ahead; their stage was last used by chunk g-1
            if (g + 2 < total_chunks) {
                if (g >= 1 && !mbar_wait(bar_cfree + 8 * st2, ((g - 1) / NAS) & 1)) failed = true;
                if (elected) load_w(st2, c2);
            }
            if (!mbar_wait(bar_wfull + 8 * st, use & 1)) failed = true;
            if (!mbar_wait(bar_aready + 8 * st, use & 1)) failed = true;
            if (c == 0 && i >= 2 && !mbar_wait(bar_accfree + 8 * acc, ((i >> 1) - 1) & 1)) failed = true;
            tc_fence_after_sync();
            const uint32_t a_base = smem_u32(smem + OFF_A) + (uint32_t)st * A_STAGE;
            const uint32_t w_base = smem_u32(smem + OFF_W) + (uint32_t)st * W_STAGE;
            const uint32_t d = tmem + (uint32_t)(acc * NMAX);
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks) {
                const uint64_t dah = make_smem_desc(a_base + (uint32_t)(2 * ks) * A_PANEL, A_PANEL, 128u);
                const uint64_t dal = make_smem_desc(a_base + A_PLANE + (uint32_t)(2 * ks) * A_PANEL, A_PANEL, 128u);
                const uint64_t dbh = make_smem_desc(w_base + (uint32_t)(2 * ks) * lbo_b, lbo_b, 128u);
                const uint64_t dbl = make_smem_desc(w_base + w_plane + (uint32_t)(2 * ks) * lbo_b, lbo_b, 128u);
                if (elected) {
                    mma_f16_ss(d, dah, dbh, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                    mma_f16_ss(d, dah, dbl, idesc, 1u);
                    mma_f16_ss(d, dal, dbh, idesc, 1u);
                }
            }
            if (elected) {
                mma_commit(bar_cfree + 8 * st);
                if (c == nchunks - 1) mma_commit(bar_accfull + 8 * acc);
            }
            __syncwarp();
            if (++c == nchunks) { c = 0; ++i; }
            if (++st == NAS) { st = 0; ++use; }
            if (++st2 == NAS) st2 = 0;
            if (++c2 == nchunks) c2 = 0;
        }
    } else if (warp >= 8) {
        // =========================================================================== producers
        const int ptid = tid - 256;
        const int q = ptid & 7, rg = ptid >> 3;              // conv: channel quad of the chunk, 8-row group
        const int xrow = ptid >> 3;                           // loads: rows xrow + 16*it, 16-byte piece q
        const uint32_t x_smem = smem_u32(smem + OFF_X);
        int* rows_s = reinterpret_cast<int*>(smem + OFF_SRC);   // GX: table row of every tile row (XROWS <= 144 ints)

        // ---- load stream (runs three chunks ahead of the conv stream) ---------------------------
        int ld_i = 0, ld_c = 0, ld_s = 0;                     // tile, chunk, x stage of the next chunk to load
        const float* ld_base = p.X;                           // row xrow of the tile, this thread's piece
        uint32_t ld_mask = 0;                                 // bit it: row xrow + 16*it exists
        auto tile_setup = [&](int i) {
            const int tile = blockIdx.x + i * gridDim.x;
            int b, t0; tile_bt(tile, b, t0);
            {
                const int tf = t0 - HALO + xrow;
                ld_base = p.X + ((long long)b * p.T + tf) * K + q * 4;
                ld_mask = 0;
                // GX: the table rows of this thread's tile rows go to shared memory (the 8 threads of a row group are
                // neighbouring lanes and write the same values; the previous tile's loads have all been issued)
                if (GX) __syncwarp();
#pragma unroll
                for (int it = 0; it < XITER; ++it) {
                    const int t = tf + 16 * it;
                    if (t >= 0 && t < p.T && xrow + 16 * it < XROWS) {
                        ld_mask |= 1u << it;
                        if (GX) rows_s[xrow + 16 * it] = __ldg(p.src + (size_t)b * p.T + t);
                    }
                }
                if (GX) __syncwarp();
            }
        };
        auto load_next = [&]() {
            const uint32_t dst = x_smem + (uint32_t)ld_s * X_STAGE + (uint32_t)ptid * 16u;
#pragma unroll
            for (int it = 0; it < XITER; ++it) {
                if (XROWS % 16 == 0 || it < XITER - 1 || xrow + 16 * it < XROWS) {
                    const float* src;
                    bool ok;
                    ok = (ld_mask >> it) & 1u;
                    if (GX) src = ok ? p.X + (size_t)rows_s[xrow + 16 * it] * K + q * 4 + ld_c * KC : p.X;
                    else    src = ok ? ld_base + (size_t)(16 * it) * K + ld_c * KC : p.X;
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
        for (int g = 0; g < 3; ++g) {                         // three chunks in flight
            if (g < total_chunks) load_next();
            cp_async_commit();
        }
        int i = 0, c = 0, sx = 0, sa = 0, use = 0;
        for (int g = 0; g < total_chunks; ++g) {
            cp_async_wait<2>();                               // this thread's share of chunk g has landed
            named_bar_sync(2, NPROD);                         // ... everybody's; chunk g-1's stage is free
            if (g + 3 < total_chunks) load_next();
            cp_async_commit();

            const float* Xc = reinterpret_cast<const float*>(smem + OFF_X + (uint32_t)sx * X_STAGE);
            uint2 ahi[8], alo[8];
            {
                float4 win[8 + 2 * HALO];
#pragma unroll
                for (int k = 0; k < 8 + 2 * HALO; ++k) win[k] = reinterpret_cast<const float4*>(Xc + (rg * 8 + k) * KC)[q];
                if (MODE == MODE_DWCONV) {
                    float4 wdw[DWK], bdw;
#pragma unroll
                    for (int t = 0; t < DWK; ++t) wdw[t] = reinterpret_cast<const float4*>(dws + t * NMAX + c * KC)[q];
                    bdw = reinterpret_cast<const float4*>(dws + DWK * NMAX + c * KC)[q];
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
                        split4(o, ahi[r], alo[r]);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 8; ++r) split4(win[r], ahi[r], alo[r]);
                }
            }
            // A stage free?  (its previous chunk, g - NAS, has been consumed by the tensor core)
            if (g >= NAS && !mbar_wait(bar_cfree + 8 * sa, (use - 1) & 1)) failed = true;
            uint8_t* a_hi = smem + OFF_A + (uint32_t)sa * A_STAGE
                            + (uint32_t)(q >> 1) * A_PANEL + (uint32_t)(rg * 8) * 16u + (uint32_t)(q & 1) * 8u;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                *reinterpret_cast<uint2*>(a_hi + r * 16) = ahi[r];
                *reinterpret_cast<uint2*>(a_hi + A_PLANE + r * 16) = alo[r];
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_aready + 8 * sa);
            if (++c == nchunks) { c = 0; ++i; }
            if (++sx == NXS) sx = 0;
            if (++sa == NAS) { sa = 0; ++use; }
        }
        cp_async_wait<0>();
    } else {
        // =========================================================================== epilogue
        const int qd = warp & 3, half = warp >> 2;
        const int rbase = qd * 32 + half * 16;
        const int t4 = lane & 3, tr = lane >> 2;
        constexpr float inv_n = 1.f / (float)N;
        const uint32_t lane_addr = (uint32_t)rbase << 16;
        const bool need_stats = p.ln_g != nullptr;
        const bool act_tanh = p.act_tanh != 0;
        const float* parc = par + 2 * t4;                       // this thread's column pair

        for (int i = 0; i < my_tiles; ++i) {
            const int tile = blockIdx.x + i * gridDim.x;
            int b, t0; tile_bt(tile, b, t0);
            const int rows_valid = min(TM2, p.T - t0);
            const int acc = i & 1;
            const int row0 = rbase + tr, row1 = row0 + 8;
            const bool ok0 = row0 < rows_valid, ok1 = row1 < rows_valid;
            float* y0 = p.Y + ((size_t)b * p.T + t0 + row0) * N + 2 * t4;
            float* y1 = y0 + 8 * N;
            const uint32_t tacc = tmem + lane_addr + (uint32_t)(acc * NMAX);
            if (p.res2) {   // pull this warp's 16 skip rows (16 KB) towards L2 while the GEMM runs
                const int pr = rbase + (lane >> 1);
                if (pr < rows_valid) {
                    const size_t srow = GS ? (size_t)__ldg(p.src + (size_t)b * p.T + t0 + pr) : (size_t)b * p.T + t0 + pr;
                    const float* sp = p.res2 + srow * N + (lane & 1) * 128;
                    prefetch_l2(sp); prefetch_l2(sp + 32); prefetch_l2(sp + 64); prefetch_l2(sp + 96);
                }
            }
            const int zero_from = p.zero_from ? p.zero_from[b] : 0x7fffffff;
            const bool z0 = (t0 + row0) >= zero_from, z1 = (t0 + row1) >= zero_from;
            const bool any_zero = __any_sync(0xffffffffu, z0 || z1);
            if (!mbar_wait(bar_accfull + 8 * acc, (i >> 1) & 1)) failed = true;
            tc_fence_after_sync();

            // ---- pass 1: bias (+ tanh), first statistics; activated values go back into the accumulator.
            //      Two register buffers: the TMEM load of step h+1 is in flight while step h is processed.
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
            {
                uint32_t rA[32], rB[32];
                tmem_ld_16x256b_x8(tacc, rA);
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    uint32_t (&r)[32] = (h & 1) ? rB : rA;
                    uint32_t (&rn)[32] = (h & 1) ? rA : rB;
                    tmem_ld_wait();
                    if (h + 1 < NH) tmem_ld_16x256b_x8(tacc + (uint32_t)((h + 1) * 64), rn);
                    const int njv = (N - h * 64) >= 64 ? 8 : (N - h * 64) / 8;      // compile-time after unrolling
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < njv) {
                            const float2 bb = *reinterpret_cast<const float2*>(parc + h * 64 + 8 * j);
                            float a, bq, cq, dq;
                            if (act_tanh) {
                                a = tanh_scaled2(fmaf(__uint_as_float(r[4 * j]), kTanhScale2, bb.x));
                                bq = tanh_scaled2(fmaf(__uint_as_float(r[4 * j + 1]), kTanhScale2, bb.y));
                                cq = tanh_scaled2(fmaf(__uint_as_float(r[4 * j + 2]), kTanhScale2, bb.x));
                                dq = tanh_scaled2(fmaf(__uint_as_float(r[4 * j + 3]), kTanhScale2, bb.y));
                            } else {
                                a = __uint_as_float(r[4 * j]) + bb.x; bq = __uint_as_float(r[4 * j + 1]) + bb.y;
                                cq = __uint_as_float(r[4 * j + 2]) + bb.x; dq = __uint_as_float(r[4 * j + 3]) + bb.y;
                            }
                            s0 += a + bq; q0 = fmaf(a, a, q0); q0 = fmaf(bq, bq, q0);
                            s1 += cq + dq; q1 = fmaf(cq, cq, q1); q1 = fmaf(dq, dq, q1);
                            r[4 * j] = __float_as_uint(a); r[4 * j + 1] = __float_as_uint(bq);
                            r[4 * j + 2] = __float_as_uint(cq); r[4 * j + 3] = __float_as_uint(dq);
                        }
                    }
                    if (need_stats) {
                        tmem_st_16x256b_x8(tacc + (uint32_t)(h * 64), r);
                    } else if (N - h * 64 >= 64) {      // no LayerNorm (mel head): the values are final
                        if (any_zero) store_frag<8, true>(y0 + h * 64, y1 + h * 64, ok0, ok1, z0, z1, r);
                        else store_frag<8, false>(y0 + h * 64, y1 + h * 64, ok0, ok1, z0, z1, r);
                    } else {
                        constexpr int NJL = (N % 64) / 8 > 0 ? (N % 64) / 8 : 8;
                        if (any_zero) store_frag<NJL, true>(y0 + h * 64, y1 + h * 64, ok0, ok1, z0, z1, r);
                        else store_frag<NJL, false>(y0 + h * 64, y1 + h * 64, ok0, ok1, z0, z1, r);
                    }
                }
            }
            if (N == NMAX && need_stats) {
                tmem_st_wait();
                float ra, na, rb, nb;
                quad_stats(s0, q0, inv_n, ra, na);
                quad_stats(s1, q1, inv_n, rb, nb);
                if (!p.res2) {
                    // ---- pass 2 (plain layer): normalise and store
                    uint32_t rA[32], rB[32];
                    tmem_ld_16x256b_x8(tacc, rA);
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        uint32_t (&r)[32] = (h & 1) ? rB : rA;
                        uint32_t (&rn)[32] = (h & 1) ? rA : rB;
                        tmem_ld_wait();
                        if (h + 1 < NH) tmem_ld_16x256b_x8(tacc + (uint32_t)((h + 1) * 64), rn);
                        norm_frag(r, parc + NMAX + h * 64, parc + 2 * NMAX + h * 64, ra, na, rb, nb);
                        if (any_zero) store_frag<8, true>(y0 + h * 64, y1 + h * 64, ok0, ok1, z0, z1, r);
                        else store_frag<8, false>(y0 + h * 64, y1 + h * 64, ok0, ok1, z0, z1, r);
                    }
                } else {
                    // ---- pass 2 (block end): normalise, add the skip row, second statistics, write back
                    s0 = q0 = s1 = q1 = 0.f;
                    const float* sp0 = p.res2 + ((size_t)b * p.T + t0 + row0) * N + 2 * t4;
                    const float* sp1 = sp0 + 8 * N;
                    if (GS) {
                        sp0 = p.res2 + (size_t)(ok0 ? __ldg(p.src + (size_t)b * p.T + t0 + row0) : 0) * N + 2 * t4;
                        sp1 = p.res2 + (size_t)(ok1 ? __ldg(p.src + (size_t)b * p.T + t0 + row1) : 0) * N + 2 * t4;
                    }
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        uint32_t r[32];
                        tmem_ld_16x256b_x8(tacc + (uint32_t)(h * 64), r);
                        float2 k0[8], k1[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            k0[j] = ok0 ? __ldg(reinterpret_cast<const float2*>(sp0 + h * 64 + 8 * j)) : make_float2(0.f, 0.f);
                            k1[j] = ok1 ? __ldg(reinterpret_cast<const float2*>(sp1 + h * 64 + 8 * j)) : make_float2(0.f, 0.f);
                        }
                        tmem_ld_wait();
                        norm_frag(r, parc + NMAX + h * 64, parc + 2 * NMAX + h * 64, ra, na, rb, nb);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float a = __uint_as_float(r[4 * j]) + k0[j].x, bq = __uint_as_float(r[4 * j + 1]) + k0[j].y;
                            const float cq = __uint_as_float(r[4 * j + 2]) + k1[j].x, dq = __uint_as_float(r[4 * j + 3]) + k1[j].y;
                            s0 += a + bq; q0 = fmaf(a, a, q0); q0 = fmaf(bq, bq, q0);
                            s1 += cq + dq; q1 = fmaf(cq, cq, q1); q1 = fmaf(dq, dq, q1);
                            r[4 * j] = __float_as_uint(a); r[4 * j + 1] = __float_as_uint(bq);
                            r[4 * j + 2] = __float_as_uint(cq); r[4 * j + 3] = __float_as_uint(dq);
                        }
                        tmem_st_16x256b_x8(tacc + (uint32_t)(h * 64), r);
                    }
                    tmem_st_wait();
                    quad_stats(s0, q0, inv_n, ra, na);
                    quad_stats(s1, q1, inv_n, rb, nb);
                    // ---- pass 3: second LayerNorm, store
                    uint32_t rA[32], rB[32];
                    tmem_ld_16x256b_x8(tacc, rA);
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        uint32_t (&r)[32] = (h & 1) ? rB : rA;
                        uint32_t (&rn)[32] = (h & 1) ? rA : rB;
                        tmem_ld_wait();
                        if (h + 1 < NH) tmem_ld_16x256b_x8(tacc + (uint32_t)((h + 1) * 64), rn);
                        norm_frag(r, parc + 3 * NMAX + h * 64, parc + 4 * NMAX + h * 64, ra, na, rb, nb);
                        if (any_zero) store_frag<8, true>(y0 + h * 64, y1 + h * 64, ok0, ok1, z0, z1, r);
                        else store_frag<8, false>(y0 + h * 64, y1 + h * 64, ok0, ok1, z0, z1, r);
                    }
                }
            }
            // the accumulator (and its scratch use) is done: the GEMM of tile i+2 may overwrite it
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accfree + 8 * acc);
        }
    }

    if (failed) atomicExch(p.err, 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int MODE, int N, bool GX = false, bool GS = false>
int launch_mode256(const Dec256Params& p, int grid, cudaStream_t s) {
    static PerDeviceSlot<bool> attr_once; bool& attr_set = attr_once.get();   // function attributes are per device
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(umma_dec256_kernel<MODE, N, GX, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    ES_CUDA(launch_pdl(umma_dec256_kernel<MODE, N, GX, GS>, grid, NTHR, SMEM_BYTES, s, p));
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace

bool umma_dec256_supported(int K, int dw_k, int N, int mode) {
    if (N != 256 && N != 80) return false;
    if (K % KC || K < KC || K > 512) return false;
    if (mode == MODE_DWCONV && (K != 256 || dw_k != DWK || N != 256)) return false;
    if (mode != MODE_DWCONV && mode != MODE_PLAIN) return false;
    return true;
}

// mode: 0 depthwise layer, 2 plain (mel head / stand-alone projection)
int launch_umma_dec256(int mode, int B, int T, int K, int N, const float* X, const float* dw_w, const float* dw_b, const void* w_chunks,
                       const float* bias, int act_tanh, const float* ln_g, const float* ln_b,
                       const float* res2, const float* ln2_g, const float* ln2_b, const int* zero_from,
                       float* Y, cudaStream_t s, const int2* tile_list, const int* tile_count) {
    ES_CHECK(w_chunks && X && Y && bias, "null tensor");
    ES_CHECK(umma_dec256_supported(K, DWK, N, mode), "shape outside the wide decoder kernel's envelope");
    ES_CHECK(!(ln_g || res2) || N == 256, "LayerNorm epilogue needs N == 256");
    ES_CHECK(!res2 || ln_g, "the skip path needs the first LayerNorm");
    int* err_flag = umma_err_flag();
    ES_CHECK(err_flag, "cannot allocate the device error flag");
    static PerDeviceSlot<int> n_sm_once; int& n_sm = n_sm_once.get();
    if (!n_sm) {
        int dev = 0;
        ES_CUDA(cudaGetDevice(&dev));
        ES_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    Dec256Params p;
    p.B = B; p.T = T; p.K = K; p.X = X;
    p.dw_w = dw_w; p.dw_b = dw_b; p.w_chunks = w_chunks; p.bias = bias; p.act_tanh = act_tanh;
    p.ln_g = ln_g; p.ln_b = ln_b; p.res2 = res2; p.ln2_g = ln2_g; p.ln2_b = ln2_b;
    p.zero_from = zero_from; p.Y = Y; p.err = err_flag; p.tile_list = tile_list; p.tile_count = tile_count; p.src = nullptr;
    const int n_tiles = B * ((T + TM2 - 1) / TM2);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    switch (mode) {
        case MODE_DWCONV: return launch_mode256<MODE_DWCONV, 256>(p, grid, s);
        default:
            if (N == 256) return launch_mode256<MODE_PLAIN, 256>(p, grid, s);
            return launch_mode256<MODE_PLAIN, 80>(p, grid, s);
    }
}

// Depthwise layer of the FIRST decoder block with the length regulator fused in: gx -- the input rows, gs -- the skip
// rows (block-end layer) are rows src[b*T + t] of the projection table X / res2 (es_umma_dec.cu has the same pair).
int launch_umma_dec256_gathered(int B, int T, const float* X, const float* dw_w, const float* dw_b, const void* w_chunks,
                                const float* bias, const float* ln_g, const float* ln_b,
                                const float* res2, const float* ln2_g, const float* ln2_b, const int* src, bool gx, bool gs,
                                float* Y, cudaStream_t s, const int2* tile_list, const int* tile_count) {
    ES_CHECK(w_chunks && X && Y && bias && ln_g && src, "null tensor");
    ES_CHECK(gx || gs, "nothing to gather");
    ES_CHECK(!gs || res2, "the gathered skip needs the table");
    int* err_flag = umma_err_flag();
    ES_CHECK(err_flag, "cannot allocate the device error flag");
    static PerDeviceSlot<int> n_sm_once; int& n_sm = n_sm_once.get();
    if (!n_sm) {
        int dev = 0;
        ES_CUDA(cudaGetDevice(&dev));
        ES_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    Dec256Params p;
    p.B = B; p.T = T; p.K = 256; p.X = X;
    p.dw_w = dw_w; p.dw_b = dw_b; p.w_chunks = w_chunks; p.bias = bias; p.act_tanh = 1;
    p.ln_g = ln_g; p.ln_b = ln_b; p.res2 = res2; p.ln2_g = ln2_g; p.ln2_b = ln2_b;
    p.zero_from = nullptr; p.Y = Y; p.err = err_flag; p.tile_list = tile_list; p.tile_count = tile_count; p.src = src;
    const int n_tiles = B * ((T + TM2 - 1) / TM2);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    if (gx && gs) return launch_mode256<MODE_DWCONV, 256, true, true>(p, grid, s);
    if (gx) return launch_mode256<MODE_DWCONV, 256, true, false>(p, grid, s);
    return launch_mode256<MODE_DWCONV, 256, false, true>(p, grid, s);
}

}  // namespace es

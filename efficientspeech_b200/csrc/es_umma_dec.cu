// Decoder-side fused tensor-core kernel (sm_100a: tcgen05 + TMEM + bulk-async copies),
// warp-specialised and software-pipelined.
//
// One persistent CTA per SM walks 64-frame tiles of [B, T, 128] activations.  Its 13 warps
// have four roles and work on DIFFERENT tiles at the same time:
//
//   issue warp 12 (one elected lane): bulk x loads into the ring, tcgen05.mma issue + commit
//   producer warps 8..11 (tile i)
//       DWCONV  the x tile (+2-frame halo) arrives by ONE bulk async copy (TMA 1-D, mbarrier
//               complete_tx) into a 3-deep ring, so two tiles (70 KB) are always in flight per
//               SM; depthwise conv k=5 + bias in registers (sliding window)
//       PLAIN   same ring without halo / conv (mel head, stand-alone projection)
//       The results are held in registers while the previous tile's GEMM still reads the A
//       operand, then split into fp16 hi/lo and stored straight into the UMMA canonical K-major
//       no-swizzle operand layout (bank-conflict free); the issue warp then launches 24 x
//       tcgen05.mma (M64 x N x K16, kind::f16: hi*hi + hi*lo + lo*hi, fp32 accumulate) and
//       commits to an mbarrier (a dedicated warp: the blocking MMA issue never stalls a producer).  Two M=64 accumulators interleave in the same 128 TMEM columns (lanes
//       {0-15,32-47,..} and {16-31,48-63,..}), so tile i+1's GEMM never waits for tile i's
//       epilogue.  The split-fp16 weights (canonical layout prepared at pack time) are
//       bulk-loaded ONCE per CTA and stay in shared memory.
//   epilogue warps 0..3 (even tiles) and 4..7 (odd tiles)
//       tcgen05.ld 16x256b: each warp owns 16 complete rows in the mma-fragment layout (4 threads
//       per row), releases the accumulator right after the load, then runs bias -> tanh ->
//       LayerNorm [-> + skip -> LayerNorm] [-> zero padded frames] in registers (2 shuffle steps
//       per statistic) and stores 32-byte row segments straight to global memory.  No
//       shared-memory staging, no CTA-wide barrier in steady state.
//
// mbarriers: bar_x[3] (x tile landed in ring slot), bar_xfree[3] (slot consumed by the producers),
// bar_aready (A operand written), bar_mma[2] (accumulator g full == A operand free),
// bar_tfree[2] (accumulator g drained by its 4 epilogue warps).  Every wait is bounded: a
// timeout raises a device flag instead of hanging the GPU.
//
// HBM traffic per layer is exactly one read of x (+ skip on block-end layers) and one write of
// y: the kernel is HBM-bound by design.
#include <stdlib.h>
#include <string.h>

#include "es_common.cuh"
#include "es_kernels.cuh"
#include "es_umma.cuh"

namespace es {
namespace {

using namespace umma;

constexpr int TM = 64;                  // frames per tile (UMMA M)
constexpr int CK = 128;                 // K = input channels
constexpr int DWK = 5;                  // depthwise taps
constexpr int NTHR = 416;               // 13 warps: 0..7 epilogue, 8..11 producer, 12 issue
constexpr int NPROD = 128;
constexpr int NSTAGE = 3;               // x ring depth
constexpr uint32_t A_LBO = 144;         // 128-byte core matrix + 16 B pad: conflict-free 8-byte lane stores
constexpr uint32_t A_SBO = 16 * A_LBO;  // 2304: one 8-row group = 16 K-chunks
constexpr uint32_t A_PLANE = (TM / 8) * A_SBO;      // 18432 bytes per fp16 plane (hi or lo)
constexpr uint32_t X_STAGE = (TM + DWK - 1) * CK * 4;   // 34816 bytes per ring slot (68 rows)

// shared memory map (dynamic, 1024-aligned base)
constexpr uint32_t OFF_XS = 0;
constexpr uint32_t OFF_A = OFF_XS + NSTAGE * X_STAGE;         // hi plane, then lo plane
constexpr uint32_t OFF_W = OFF_A + 2 * A_PLANE;               // W hi [K/8][N][8], then W lo
constexpr uint32_t W_PLANE_MAX = 128 * CK * 2;                // 32768
constexpr uint32_t OFF_PAR = OFF_W + 2 * W_PLANE_MAX;         // bias, ln g/b, ln2 g/b: 5 x 128 floats
constexpr uint32_t OFF_DW = OFF_PAR + 5 * 128 * 4;            // depthwise taps + bias: 6 x 128 floats
constexpr int MAP_LD = 72;                                    // row-map stride per ring slot (68 tile rows)
constexpr uint32_t OFF_MAP = OFF_DW + 6 * 128 * 4;            // gathered x tiles: tile row -> staged row, per ring slot
constexpr uint32_t OFF_BAR = OFF_MAP + NSTAGE * MAP_LD * 4;   // 12 mbarriers + tmem base
constexpr int TL_CACHE = 64;                                  // ragged schedule: this CTA's first 64 tile coordinates, staged once
constexpr uint32_t OFF_TL = OFF_BAR + 128;
constexpr uint32_t SMEM_BYTES = OFF_TL + TL_CACHE * 8;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

enum { MODE_DWCONV = 0, MODE_PLAIN = 2 };

struct UmmaDecParams {
    int B, T, N;                 // N = output channels (128, or 80 for the mel head)
    const float* X;              // [B,T,128] (GX: the projection table)
    const float* dw_w;           // [5][128]
    const float* dw_b;           // [128]
    const void* w_h16;           // canonical split-fp16 weights: [2][16][N][8] halves
    const float* bias;           // [N]
    int act_tanh;
    const float* ln_g; const float* ln_b;       // LayerNorm over N, or null
    const float* res2; const float* ln2_g; const float* ln2_b;   // out = LN2(out + res2), or null
    const int* zero_from;        // [B] rows t >= zero_from[b] zeroed, or null
    const int* src;              // GX / GS kernels: frame -> table row map [B*T] (es_gather.cu); X / res2 is the table
    int pad_id;                  // GX: table row of the zero-padded frames (the largest row index)
    const int2* tile_list;       // ragged scheduling: the (b, t0) of the tiles that can reach a valid frame, or null = all tiles
    const int* tile_count;       //                    their number (device memory, written by tile_list_kernel)
    float* Y;                    // [B,T,N]
    int* err;                    // device error flag (mbarrier timeout)
    long long* trace;            // debug: per-role clock64 stamps of CTA 0 ([4 roles][32 tiles][8 events]) or null
};

// tanh(x) = 1 - 2 / (1 + e^{2x}) with e^{2x} = 2^{x * 2 log2(e)}: two MUFU ops (ex2, rcp) and two FMAs.
// The argument arrives pre-scaled (acc * c + bias * c, c = 2 log2 e).  Saturates correctly at
// +-inf (ex2 -> inf -> rcp -> 0 -> 1; ex2 -> 0 -> rcp(1) -> -1); absolute error ~1.2e-7.
constexpr float kTanhScale = 2.8853900817779268f;
__device__ __forceinline__ float ex2_approx(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
    return e;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// byte offset of channels 4*lane..4*lane+3 of tile row `row` inside an A plane
__device__ __forceinline__ uint32_t a_off(int row, int lane) {
    return (uint32_t)(lane >> 1) * A_LBO + (uint32_t)(row >> 3) * A_SBO + (uint32_t)(row & 7) * 16u +
           (uint32_t)(lane & 1) * 8u;
}

// debug time stamps (CTA 0 only, one lane per role); compiled in, one predictable branch when off
#define ES_TRACE(role, iter, ev)                                                             \
    do {                                                                                     \
        if (p.trace && blockIdx.x == 0 && (iter) < 32) p.trace[((role) * 32 + (iter)) * 8 + (ev)] = clock64(); \
    } while (0)

// Epilogue math of one warp's 16 x 128 accumulator slab, in registers (fragment layout of
// tcgen05.ld.16x256b: r[4j + 2*row + b] = column 8j + 2*t4 + b of this thread's row `row`):
// bias -> tanh -> LayerNorm [-> + skip -> LayerNorm] [-> zero padded frames] -> 16-byte stores.
// `res0` / `res1` point at the skip rows of this thread's two rows, `yrow0` at its first output row (the second
// is 8 rows further); zero_rows: rows >= it are zeroed.
template <bool RES2, typename Stamp>
__device__ __forceinline__ void epilogue_math(uint32_t (&r)[64], const float* par, int N, int nj, float inv_n,
                                              bool act_tanh, bool has_ln, const float* res0, const float* res1,
                                              float* yrow0, bool ok0, bool ok1, int zero_rows, int t4, Stamp stamp) {
    auto ld_skip = [](const float* ptr) { return __ldg(reinterpret_cast<const ulonglong2*>(ptr)); };
    // block-end layers: the 32 skip values of this thread's first row are requested NOW (L2-prefetched
    // above) and land while the tanh / LayerNorm math below runs; the second row's follow while the
    // first row is normalised (keeps the live set at 64 + 32 registers)
    // (16-byte loads of 4 consecutive columns -- 8 rows x 64 B per warp-wide load -- un-swapped into
    // the fragment layout by one quad shuffle per pair when they are consumed, see the stores below)
    ulonglong2 sk[8];
    const bool odd = t4 & 1;
    const int qcol = odd ? 8 + 2 * (t4 - 1) : 2 * t4;   // first of this thread's 4 consecutive columns
    const float* s0 = res0 + qcol;
    const float* s1 = res1 + qcol;
    if (RES2) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
            sk[k] = ok0 ? ld_skip(s0 + 16 * k) : make_ulonglong2(0ull, 0ull);
    }
    // v[2j + row]: columns 8j + 2*t4, +1 of this thread's row `row`, as one packed fp32 pair
    f32x2 v[32];
    if (act_tanh) {
        // tanh(x) = 1 - 2 / (1 + 2^(x * 2 log2 e)): FFMA2, 2 x MUFU.EX2, FADD2, 2 x MUFU.RCP, FFMA2 per pair
        const f32x2 cs = pk2(kTanhScale, kTanhScale), one2 = pk2(1.f, 1.f), mtwo2 = pk2(-2.f, -2.f);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const f32x2 bb = *reinterpret_cast<const f32x2*>(par + 8 * j + 2 * t4);
#pragma unroll
            for (int row = 0; row < 2; ++row) {
#ifdef ES_EXP_NOTANH      // timing experiment (tools/gpu_exp.sh): never defined in the product build
                v[2 * j + row] = fma2(pk2u(r[4 * j + 2 * row], r[4 * j + 2 * row + 1]), cs, bb);
                (void)one2; (void)mtwo2;
#else
                const float2 a = up2(fma2(pk2u(r[4 * j + 2 * row], r[4 * j + 2 * row + 1]), cs, bb));
                const float2 d = up2(add2(pk2(ex2_approx(a.x), ex2_approx(a.y)), one2));
                v[2 * j + row] = fma2(mtwo2, pk2(rcp_approx(d.x), rcp_approx(d.y)), one2);
#endif
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const f32x2 bb = *reinterpret_cast<const f32x2*>(par + 8 * j + 2 * t4);
            v[2 * j] = add2(pk2u(r[4 * j], r[4 * j + 1]), bb);
            v[2 * j + 1] = add2(pk2u(r[4 * j + 2], r[4 * j + 3]), bb);
        }
    }
    if (has_ln) fragment_layernorm2_p(v, par + 128, par + 256, t4, inv_n);   // LayerNorm only with N == 128
    stamp(3);
    if (RES2) {
        // thread pairs (t4, t4^1) hold [group 2k: own pair, partner's pair] (even t4) or
        // [group 2k+1: partner's pair, own pair] (odd t4): swap the foreign halves
        auto add_skip = [&](int row) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const f32x2 recv = __shfl_xor_sync(0xffffffffu, odd ? sk[k].x : sk[k].y, 1);
                v[4 * k + row] = add2(v[4 * k + row], odd ? recv : sk[k].x);
                v[4 * k + 2 + row] = add2(v[4 * k + 2 + row], odd ? sk[k].y : recv);
            }
        };
        add_skip(0);
#pragma unroll
        for (int k = 0; k < 8; ++k)
            sk[k] = ok1 ? ld_skip(s1 + 16 * k) : make_ulonglong2(0ull, 0ull);
        fragment_layernorm_row_p<0>(v, par + 384, par + 512, t4, inv_n);
        add_skip(1);
        fragment_layernorm_row_p<1>(v, par + 384, par + 512, t4, inv_n);
    }
    stamp(4);
    if (zero_rows <= 8) {       // mel head: frames past mel_len are zeroed (networks.py:424-427); rare rows
        if (zero_rows <= 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[2 * j] = 0ull;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[2 * j + 1] = 0ull;
    }
    // Stores.  In the fragment layout a warp-wide 8-byte store touches 8 rows x 32 B: 8 L1 wavefronts
    // for 256 B, and the L1 data pipe (shared with the operand fetches of the tensor core) is the
    // busiest unit of this kernel.  Neighbouring threads of a quad therefore swap one column pair per
    // two 8-column groups, so that every thread owns 4 consecutive columns and a warp-wide 16-byte
    // store covers 8 rows x 64 B: half the wavefronts per byte.
    float* y0 = yrow0 + qcol;
    float* y1 = y0 + 8 * N;
#ifdef ES_EXP_NOSTORE     // timing experiment: keep the math alive, store (practically) nothing
    {
        f32x2 acc = 0ull;
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= v[j];
        if (acc == 0x7fc123457fc12345ull && ok0) *reinterpret_cast<f32x2*>(y0) = acc;
        return;
    }
#endif
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (2 * k < nj) {                                // nj is even (N % 16 == 0)
#pragma unroll
            for (int row = 0; row < 2; ++row) {
                const f32x2 keep = odd ? v[4 * k + 2 + row] : v[4 * k + row];
                const f32x2 send = odd ? v[4 * k + row] : v[4 * k + 2 + row];
                const f32x2 recv = __shfl_xor_sync(0xffffffffu, send, 1);
                ulonglong2 o;
                o.x = odd ? recv : keep;
                o.y = odd ? keep : recv;
                if (row == 0 ? ok0 : ok1) *reinterpret_cast<ulonglong2*>((row == 0 ? y0 : y1) + 16 * k) = o;
            }
        }
    }
}

// RES2: block-end layer (skip add + second LayerNorm) -- a compile-time switch, so that the plain layers do not
// reserve the 32 skip registers and ptxas can hoist the parameter loads of the epilogue instead
// GX / GS (DWCONV only): the input rows (GX) and / or the skip rows (GS) are rows of an L2-resident table
// addressed through the frame -> row map p.src -- the length regulator fused into the first decoder block.
// GX: the rows of one utterance are nondecreasing in t and CONTIGUOUS in the table, so the distinct rows a
// tile needs arrive as ONE bulk copy (~T/N times fewer bytes than the tile), plus the padded-frame row; the
// issue warp writes a tile-row -> staged-row map next to it and the producers read their conv window through
// that map (rows outside the utterance map to -1 = zeros, so no zero fill either).  Runs of zero-duration
// phonemes that would overflow the slot fall back to one 512-byte copy per frame.  The row indices are
// fetched one iteration ahead.  GS: the epilogue reads its skip rows through the map.
template <int MODE, bool RES2, bool GX = false, bool GS = false>
__global__ void __launch_bounds__(NTHR, 1)
umma_dec_kernel(const UmmaDecParams p) {
    static_assert(!(GX || GS) || MODE == MODE_DWCONV, "gathered rows: depthwise layers only");
    static_assert(!GS || RES2, "gathered skip rows need the block-end epilogue");
    constexpr int HALO = (MODE == MODE_DWCONV) ? DWK / 2 : 0;
    constexpr int XROWS = TM + 2 * HALO;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* a_hi = smem + OFF_A;
    float* par = reinterpret_cast<float*>(smem + OFF_PAR);
    int* xmap = reinterpret_cast<int*>(smem + OFF_MAP);
    const uint32_t bar_x = smem_u32(smem + OFF_BAR);          // [3] x tile landed
    const uint32_t bar_w = bar_x + 24;                        //     weights landed
    const uint32_t bar_mma = bar_x + 32;                      // [2] accumulator g full / A operand free
    const uint32_t bar_tfree = bar_x + 48;                    // [2] accumulator g drained
    const uint32_t bar_xfree = bar_x + 64;                    // [3] ring slot consumed
    const uint32_t bar_aready = bar_x + 88;                   //     A operand written by the 4 producer warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 96);

    const int N = p.N;
    const int tiles_per_utt = (p.T + TM - 1) / TM;
    int n_tiles = p.B * tiles_per_utt;
    // tile index -> (utterance, first frame): dense order, or the compacted list of the ragged schedule (es_gather.cu)
    const int2* s_tiles = reinterpret_cast<const int2*>(smem + OFF_TL);
    bool use_list = false;        // set after the dependency wait: a list that holds EVERY tile is the dense order itself
    auto tile_bt = [&](int tile, int& b, int& t0) {
        if (use_list) {
            // (a dependent global load per tile and role costs more than skipping the tiles saves: the CTA's own
            // entries are staged in shared memory right after the dependency wait)
            const int k = (tile - (int)blockIdx.x) / (int)gridDim.x;
            const int2 v = k < TL_CACHE ? s_tiles[k] : __ldg(p.tile_list + tile);
            b = v.x; t0 = v.y;
        } else {
            b = tile / tiles_per_utt; t0 = (tile - b * tiles_per_utt) * TM;
        }
    };
    const uint32_t w_plane = (uint32_t)N * CK * 2u;

    // ---- one-time setup ---------------------------------------------------------------------
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 128);      // 128 fp32 columns x 128 lanes = two M64 accumulators
    if (tid == 0) {
        for (int k = 0; k < NSTAGE; ++k) mbar_init(bar_x + 8 * k, 1);
        mbar_init(bar_w, 1);
        mbar_init(bar_mma, 1);
        mbar_init(bar_mma + 8, 1);
        mbar_init(bar_tfree, 4);
        mbar_init(bar_tfree + 8, 4);
        for (int k = 0; k < NSTAGE; ++k) mbar_init(bar_xfree + 8 * k, 4);
        mbar_init(bar_aready, 4);
        fence_mbar_init();
        // the weights do not depend on the predecessor kernel: their bulk load starts before pdl_wait()
        mbar_arrive_expect_tx(bar_w, 2 * w_plane);
        bulk_g2s(smem_u32(smem + OFF_W), p.w_h16, w_plane, bar_w);
        bulk_g2s(smem_u32(smem + OFF_W) + w_plane, reinterpret_cast<const uint8_t*>(p.w_h16) + w_plane, w_plane, bar_w);
    }
    for (int i = tid; i < 128; i += NTHR) {
        par[i] = (i < N) ? __ldg(p.bias + i) * (p.act_tanh ? kTanhScale : 1.f) : 0.f;   // pre-scaled for tanh
        par[128 + i] = (p.ln_g && i < N) ? __ldg(p.ln_g + i) : 0.f;
        par[256 + i] = (p.ln_g && i < N) ? __ldg(p.ln_b + i) : 0.f;
        par[384 + i] = (p.ln2_g && i < N) ? __ldg(p.ln2_g + i) : 0.f;
        par[512 + i] = (p.ln2_g && i < N) ? __ldg(p.ln2_b + i) : 0.f;
    }
    if (MODE == MODE_DWCONV) {
        float* dws = reinterpret_cast<float*>(smem + OFF_DW);
        for (int i = tid; i < (DWK + 1) * CK; i += NTHR) dws[i] = i < DWK * CK ? __ldg(p.dw_w + i) : __ldg(p.dw_b + i - DWK * CK);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    bool failed = false;
    pdl_launch_dependents();      // the next kernel may start its prologue
    pdl_wait();                   // the previous kernel's output is complete and visible from here on
    if (p.tile_count) {
        n_tiles = *reinterpret_cast<const volatile int*>(p.tile_count);
        use_list = n_tiles != p.B * tiles_per_utt;           // CTA-uniform
        if (use_list) {
            if (tid < TL_CACHE) {
                const int tile = blockIdx.x + tid * gridDim.x;
                if (tile < n_tiles) reinterpret_cast<int2*>(smem + OFF_TL)[tid] = __ldg(p.tile_list + tile);
            }
            __syncthreads();
        }
    }

    if (warp == 12) {
        // =========================================================================== issue warp
        // rows [t0-HALO, t0+TM+HALO) clipped to the utterance -> ring slot, one bulk copy
        auto issue_x = [&](int tile, int slot) {
            int b, t0; tile_bt(tile, b, t0);
            const int lo = max(t0 - HALO, 0), hi = min(t0 + TM + HALO, p.T);
            const uint32_t bytes = (uint32_t)(hi - lo) * CK * 4u;
            mbar_arrive_expect_tx(bar_x + 8 * slot, bytes);
            bulk_g2s(smem_u32(smem + OFF_XS) + (uint32_t)slot * X_STAGE + (uint32_t)(lo - (t0 - HALO)) * CK * 4u,
                     p.X + ((size_t)b * p.T + lo) * CK, bytes, bar_x + 8 * slot);
        };
        const uint32_t idesc = make_idesc_f16(TM, N);
        const uint32_t lbo_b = (uint32_t)N * 16u;
        // UMMA shared-memory descriptors of K step 0 (A: padded canonical tile; B: resident weights);
        // they advance by a constant per K step (start-address field, 16-byte units)
        const uint64_t dah0 = make_smem_desc(smem_u32(a_hi), A_LBO, A_SBO);
        const uint64_t dal0 = make_smem_desc(smem_u32(a_hi) + A_PLANE, A_LBO, A_SBO);
        const uint64_t dbh0 = make_smem_desc(smem_u32(smem + OFF_W), lbo_b, 128u);
        const uint64_t dbl0 = make_smem_desc(smem_u32(smem + OFF_W) + w_plane, lbo_b, 128u);
        const uint64_t db_step = (uint64_t)((2u * lbo_b) >> 4);
        const bool elected = elect_one();                    // one lane issues on behalf of the CTA; the
                                                             // control flow stays warp-uniform
        // gathered x: lane l owns frames lo + l, lo + l + 32, lo + l + 64 of a tile (<= 68 frames)
        constexpr bool gx = GX;
        int sx[3] = {-1, -1, -1};
        auto load_src = [&](int tile) {
            int b, t0; tile_bt(tile, b, t0);
            const int lo = max(t0 - HALO, 0), hi = min(t0 + TM + HALO, p.T);
            const int* sp = p.src + (size_t)b * p.T;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int r = lo + lane + 32 * k;
                sx[k] = r < hi ? __ldg(sp + r) : -1;
            }
        };
        auto issue_rows = [&](int tile, int slot) {
            int b, t0; tile_bt(tile, b, t0);
            const int lo = max(t0 - HALO, 0), hi = min(t0 + TM + HALO, p.T);
            const int head = lo - (t0 - HALO), nfr = hi - lo;   // tile rows [head, head + nfr) are frames lo..hi-1
            const int first = __shfl_sync(0xffffffffu, sx[0], 0);
            int mx = -1;
            bool anypad = false;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (sx[k] == p.pad_id) anypad = true;
                else mx = max(mx, sx[k]);                    // -1 (no frame) never wins
            }
            mx = __reduce_max_sync(0xffffffffu, mx);
            const int has_pad = __any_sync(0xffffffffu, anypad) ? 1 : 0;
            const int nr = mx >= 0 ? mx - first + 1 : 0;     // distinct table rows [first, mx]; padded frames are the tail
            const bool compact = nr + has_pad <= XROWS;
            int* mp = xmap + slot * MAP_LD;
            for (int r = lane; r < XROWS; r += 32)
                if (r < head || r >= head + nfr) mp[r] = -1;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int r = head + lane + 32 * k;
                if (sx[k] >= 0) mp[r] = compact ? (sx[k] == p.pad_id ? nr : sx[k] - first) : r;
            }
            __syncwarp();                                    // the map is written before the barrier can complete
            const uint32_t dst = smem_u32(smem + OFF_XS) + (uint32_t)slot * X_STAGE;
            if (compact) {
                if (elected) {
                    mbar_arrive_expect_tx(bar_x + 8 * slot, (uint32_t)(nr + has_pad) * CK * 4u);
                    if (nr) bulk_g2s(dst, p.X + (size_t)first * CK, (uint32_t)nr * CK * 4u, bar_x + 8 * slot);
                    if (has_pad) bulk_g2s(dst + (uint32_t)nr * CK * 4u, p.X + (size_t)p.pad_id * CK, CK * 4u, bar_x + 8 * slot);
                }
            } else {
                if (elected) mbar_arrive_expect_tx(bar_x + 8 * slot, (uint32_t)nfr * CK * 4u);
                __syncwarp();                                // the byte count is registered before any copy can complete
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    if (sx[k] >= 0) bulk_g2s(dst + (uint32_t)(head + lane + 32 * k) * CK * 4u, p.X + (size_t)sx[k] * CK, CK * 4u, bar_x + 8 * slot);
            }
        };
        for (int k = 0; k < NSTAGE; ++k) {
            const int tile = blockIdx.x + k * gridDim.x;
            if (tile < n_tiles) {
                if (gx) { load_src(tile); issue_rows(tile, k); }
                else if (elected) issue_x(tile, k);
            }
        }
        if (!mbar_wait(bar_w, 0)) failed = true;
        __syncwarp();
        int i = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
            const int g = i & 1, u = i >> 1, slot = i % NSTAGE;
            const int next = tile + NSTAGE * (int)gridDim.x;
            if (gx && next < n_tiles) load_src(next);                    // row indices land while this tile's GEMM is set up
            if (elected) ES_TRACE(0, i, 0);
            if (!mbar_wait(bar_aready, i & 1)) failed = true;            // A operand of tile i written
            if (elected) ES_TRACE(0, i, 1);
            // accumulator g drained by its epilogue warps (its previous use was tile i-2)?
            if (u > 0 && !mbar_wait(bar_tfree + 8 * g, (u - 1) & 1)) failed = true;
            tc_fence_after_sync();
            if (elected) ES_TRACE(0, i, 2);
            const uint32_t acc = tmem + ((uint32_t)(16 * g) << 16);    // M64 accumulators interleave by 16 lanes
#pragma unroll
            for (int k = 0; k < CK / 16; ++k) {
                const uint64_t da = (uint64_t)((uint32_t)(2 * k) * A_LBO >> 4);
                const uint64_t db = (uint64_t)k * db_step;
                if (elected) {
                    mma_f16_ss(acc, dah0 + da, dbh0 + db, idesc, k > 0 ? 1u : 0u);
                    mma_f16_ss(acc, dah0 + da, dbl0 + db, idesc, 1u);
                    mma_f16_ss(acc, dal0 + da, dbh0 + db, idesc, 1u);
                }
            }
            if (elected) { mma_commit(bar_mma + 8 * g); ES_TRACE(0, i, 3); }
            // ring slot of tile i was released by the producers before they signalled bar_aready:
            // the tile three steps ahead starts streaming into it
            if (next < n_tiles) {
                if (!mbar_wait(bar_xfree + 8 * slot, (i / NSTAGE) & 1)) failed = true;
                if (gx) issue_rows(next, slot);
                else if (elected) issue_x(next, slot);
            }
            if (elected) ES_TRACE(0, i, 4);
            __syncwarp();
        }
    } else if (warp >= 8) {
        // =========================================================================== producers
        const int pw = warp - 8, ptid = tid - 256;
        const ulonglong2* dwp = reinterpret_cast<const ulonglong2*>(smem + OFF_DW);   // [5 taps + bias][32 lanes] 4 channels as 2 fp32 pairs

        int i = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
            int b, t0; tile_bt(tile, b, t0);
            const int slot = i % NSTAGE;
            float* Xs = reinterpret_cast<float*>(smem + OFF_XS + (uint32_t)slot * X_STAGE);
            const bool tr_on = (pw == 0 && lane == 0);
            if (tr_on) ES_TRACE(1, i, 0);

            if (GX) {
                if (!mbar_wait(bar_x + 8 * slot, (i / NSTAGE) & 1)) failed = true;   // staged rows + row map
                if (tr_on) ES_TRACE(1, i, 1);
            } else {
                // zero the halo / tail rows the bulk copy does not cover (utterance boundaries only)
                const int lo = max(t0 - HALO, 0), hi = min(t0 + TM + HALO, p.T);
                const int head = lo - (t0 - HALO), tail0 = hi - (t0 - HALO);
                const bool edge = head > 0 || tail0 < XROWS;
                if (edge) {
                    for (int k = ptid; k < (head + XROWS - tail0) * (CK / 4); k += NPROD) {
                        int r = k / (CK / 4);
                        const int c4 = k - r * (CK / 4);
                        if (r >= head) r = tail0 + (r - head);
                        reinterpret_cast<float4*>(Xs + r * CK)[c4] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                if (!mbar_wait(bar_x + 8 * slot, (i / NSTAGE) & 1)) failed = true;
                if (edge) named_bar_sync(1, NPROD);          // zero fill visible to every producer warp
                if (tr_on) ES_TRACE(1, i, 1);
            }
            // warp pw -> tile rows 16pw..16pw+15 (two 8-row core-matrix groups).  Both groups are computed
            // into registers BEFORE waiting for the A operand to be released, so that after the previous
            // GEMM completes only the 32 shared-memory stores remain on the critical path.
            uint2 ahi[16], alo[16];                          // this lane's share: 16 rows x 4 channels, split fp16
            {
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {
                    const int r0 = pw * 16 + pass * 8;
                    // per-lane depthwise taps for channels 4*lane..4*lane+3 (3 KB in shared memory; kept out of
                    // the persistent register set so the staged A rows fit without spilling); channel pairs
                    // are packed fp32x2 operands: 10 FFMA2 per row instead of 20 FFMA
                    ulonglong2 wdw[DWK], bdw;
                    if (MODE == MODE_DWCONV) {
#pragma unroll
                        for (int t = 0; t < DWK; ++t) wdw[t] = dwp[t * 32 + lane];
                        bdw = dwp[DWK * 32 + lane];
                    }
                    ulonglong2 win[8 + 2 * HALO];
                    if (GX) {
                        const int* mp = xmap + slot * MAP_LD + r0;
#pragma unroll
                        for (int k = 0; k < 8 + 2 * HALO; ++k) {
                            const int mrow = mp[k];
                            win[k] = mrow >= 0 ? reinterpret_cast<const ulonglong2*>(Xs + mrow * CK)[lane] : make_ulonglong2(0ull, 0ull);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 8 + 2 * HALO; ++k) win[k] = reinterpret_cast<const ulonglong2*>(Xs + (r0 + k) * CK)[lane];
                    }
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        ulonglong2 o;
#ifdef ES_EXP_NODW         // timing experiment: no depthwise FMAs
                        if (false) {
#else
                        if (MODE == MODE_DWCONV) {
#endif
                            o = bdw;
#pragma unroll
                            for (int t = 0; t < DWK; ++t) {
                                o.x = fma2(wdw[t].x, win[r + t].x, o.x);
                                o.y = fma2(wdw[t].y, win[r + t].y, o.y);
                            }
                        } else {
                            o = win[r + HALO];
                        }
#ifdef ES_EXP_NOSPLIT      // timing experiment: no fp16 split (raw bits stored)
                        ahi[pass * 8 + r] = make_uint2((uint32_t)o.x, (uint32_t)o.y);
                        alo[pass * 8 + r] = make_uint2((uint32_t)(o.x >> 32), (uint32_t)(o.y >> 32));
#else
                        split4(o, ahi[pass * 8 + r], alo[pass * 8 + r]);
#endif
                    }
                }
            }
            // this warp has read everything it needs from ring slot `slot` (and from srcs)
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_xfree + 8 * slot);
            if (tr_on) ES_TRACE(1, i, 2);
            // A operand free?  (the GEMM of the previous tile has read it)
            if (i > 0 && !mbar_wait(bar_mma + 8 * ((i - 1) & 1), ((i - 1) >> 1) & 1)) failed = true;
            if (tr_on) ES_TRACE(1, i, 3);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const uint32_t off = a_off(pw * 16 + r, lane);
                *reinterpret_cast<uint2*>(a_hi + off) = ahi[r];
                *reinterpret_cast<uint2*>(a_hi + A_PLANE + off) = alo[r];
            }
            if (tr_on) ES_TRACE(1, i, 4);
            fence_proxy_async_smem();                  // generic-proxy stores -> visible to the tensor core
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_aready);
            if (tr_on) ES_TRACE(1, i, 5);
        }
    } else {
        // =========================================================================== epilogue
        const int q = warp & 3, g = warp >> 2;                // TMEM lane quarter; tile parity served by this warp
        const int rbase = q * 16;                             // tile rows 16q..16q+15 live in lanes 32q+16g..+15
        const int t4 = lane & 3, tr = lane >> 2;
        const float inv_n = 1.f / (float)N;
        const int nj = N >> 3;                                // 8-column groups

        int i = g;
        for (int tile = blockIdx.x + g * gridDim.x; tile < n_tiles; tile += 2 * gridDim.x, i += 2) {
            int b, t0; tile_bt(tile, b, t0);
            const int rows_valid = min(TM, p.T - t0);
            const int u = i >> 1;
            const int row0 = rbase + tr, row1 = row0 + 8;
            const size_t g0 = (size_t)b * p.T + t0 + row0;
            const bool ok0 = row0 < rows_valid, ok1 = row1 < rows_valid;

            const bool tr_on = (q == 0 && lane == 0);
            if (tr_on) ES_TRACE(2 + g, u, 0);
            const float* res0 = nullptr;
            const float* res1 = nullptr;
            if (RES2) {
                const int pr = rbase + (lane >> 1);          // lanes 2k, 2k+1 <-> row rbase + k of this warp's 16
                if (GS) {
                    // skip rows come from the (L2-resident) table: fetch the row indices while the GEMM runs
                    const int s_pr = pr < rows_valid ? __ldg(p.src + (size_t)b * p.T + t0 + pr) : 0;
                    res0 = p.res2 + (size_t)__shfl_sync(0xffffffffu, s_pr, 2 * tr) * N;
                    res1 = p.res2 + (size_t)__shfl_sync(0xffffffffu, s_pr, 2 * tr + 16) * N;
                } else {
                    // pull this warp's 16 skip rows (8 KB) towards L2 while the GEMM runs
                    if (pr < rows_valid) {
                        const float* sp = p.res2 + ((size_t)b * p.T + t0 + pr) * N + (lane & 1) * 64;
                        prefetch_l2(sp);
                        prefetch_l2(sp + 32);
                    }
                    res0 = p.res2 + g0 * N;
                    res1 = res0 + 8 * N;
                }
            }
            if (!mbar_wait(bar_mma + 8 * g, u & 1)) failed = true;
            tc_fence_after_sync();
            if (tr_on) ES_TRACE(2 + g, u, 1);
            uint32_t r[64];
            tmem_ld_16x256b_x16(tmem + ((uint32_t)(32 * q + 16 * g) << 16), r);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tfree + 8 * g);     // accumulator drained: the GEMM of tile i+2 may start
            if (tr_on) ES_TRACE(2 + g, u, 2);
#ifdef ES_EXP_NOEPI       // timing experiment: the epilogue only drains the accumulator
            {
                uint32_t acc = 0;
#pragma unroll
                for (int j = 0; j < 64; ++j) acc ^= r[j];
                if (acc == 0x7fc12345u && ok0) p.Y[g0 * N] = __uint_as_float(acc);
                continue;
            }
#endif
            epilogue_math<RES2>(r, par, N, nj, inv_n, p.act_tanh != 0, p.ln_g != nullptr,
                                res0, res1, p.Y + g0 * N, ok0, ok1,
                                p.zero_from ? p.zero_from[b] - (t0 + row0) : 0x7fffffff, t4,
                                [&](int ev) { if (tr_on) ES_TRACE(2 + g, u, ev); });
            if (tr_on) ES_TRACE(2 + g, u, 5);
        }
    }

    if (failed) atomicExch(p.err, 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

PerDeviceSlot<int*> g_err_flags;       // one flag per device (the kernels of device d write device d's flag)
long long* g_trace = nullptr;
int g_trace_pick = 0, g_trace_count = 0;   // which launch after es_debug_set_trace is stamped (env ES_TRACE_LAUNCH)

template <int MODE, bool RES2, bool GX = false, bool GS = false>
int launch_mode(const UmmaDecParams& p, int grid, cudaStream_t s) {
    static PerDeviceSlot<bool> attr_once; bool& attr_set = attr_once.get();   // function attributes are per device
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(umma_dec_kernel<MODE, RES2, GX, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    ES_CUDA(launch_pdl(umma_dec_kernel<MODE, RES2, GX, GS>, grid, NTHR, SMEM_BYTES, s, p));
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace

bool umma_dec_supported(int C, int dw_k, int N) {
    return C == CK && dw_k == DWK && (N == 128 || N == 80);
}

// mode: 0 depthwise layer, 2 plain (mel head, per-phoneme projection)
int launch_umma_dec(int mode, int B, int T, int N, const float* X, const float* dw_w, const float* dw_b, const void* w_h16,
                    const float* bias, int act_tanh, const float* ln_g, const float* ln_b,
                    const float* res2, const float* ln2_g, const float* ln2_b, const int* zero_from,
                    float* Y, cudaStream_t s, const int2* tile_list, const int* tile_count) {
    ES_CHECK(w_h16 && X && Y && bias, "null tensor");
    ES_CHECK(N % 16 == 0 && N >= 32 && N <= 128, "N must be a multiple of 16 in [32,128]");
    ES_CHECK(!(ln_g || res2) || N == 128, "LayerNorm epilogue needs N == 128");
    int* const g_err_flag = umma_err_flag();
    ES_CHECK(g_err_flag, "cannot allocate the device error flag");
    static PerDeviceSlot<int> n_sm_once; int& n_sm = n_sm_once.get();
    if (!n_sm) {
        int dev = 0;
        ES_CUDA(cudaGetDevice(&dev));
        ES_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    UmmaDecParams p;
    p.B = B; p.T = T; p.N = N; p.X = X;
    p.dw_w = dw_w; p.dw_b = dw_b; p.w_h16 = w_h16; p.bias = bias; p.act_tanh = act_tanh;
    p.ln_g = ln_g; p.ln_b = ln_b; p.res2 = res2; p.ln2_g = ln2_g; p.ln2_b = ln2_b;
    p.zero_from = zero_from; p.src = nullptr; p.pad_id = 0; p.tile_list = tile_list; p.tile_count = tile_count;
    p.Y = Y; p.err = g_err_flag; p.trace = (g_trace && g_trace_count++ == g_trace_pick) ? g_trace : nullptr;
    const int n_tiles = B * ((T + TM - 1) / TM);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    switch (mode) {
        case MODE_DWCONV: return res2 ? launch_mode<MODE_DWCONV, true>(p, grid, s) : launch_mode<MODE_DWCONV, false>(p, grid, s);
        case MODE_PLAIN: ES_CHECK(!res2, "plain mode has no skip input"); return launch_mode<MODE_PLAIN, false>(p, grid, s);
        default: ES_CHECK(false, "unknown mode"); return 1;
    }
}

// Depthwise layer with table-gathered input and / or skip rows (the length regulator fused into the first block).
int launch_umma_dec_gathered(int B, int T, int N, const float* X, const float* dw_w, const float* dw_b,
                             const void* w_h16, const float* bias, int act_tanh, const float* ln_g, const float* ln_b,
                             const float* res2, const float* ln2_g, const float* ln2_b, const int* src, int pad_id,
                             int gather_x, int gather_res2, float* Y, cudaStream_t s, const int2* tile_list,
                             const int* tile_count) {
    ES_CHECK(w_h16 && X && Y && bias && dw_w && dw_b && src, "null tensor");
    ES_CHECK(N == 128, "gathered layers are full-width (N == 128)");
    ES_CHECK(gather_x || gather_res2, "nothing to gather");
    ES_CHECK(!gather_res2 || res2, "gathered skip without a table");
    ES_CHECK(!(gather_x && res2) || gather_res2, "a layer that gathers its input also gathers its skip (first block)");
    int* const g_err_flag = umma_err_flag();
    ES_CHECK(g_err_flag, "cannot allocate the device error flag");
    int dev = 0, n_sm = 0;
    ES_CUDA(cudaGetDevice(&dev));
    ES_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    UmmaDecParams p;
    p.B = B; p.T = T; p.N = N; p.X = X;
    p.dw_w = dw_w; p.dw_b = dw_b; p.w_h16 = w_h16; p.bias = bias; p.act_tanh = act_tanh;
    p.ln_g = ln_g; p.ln_b = ln_b; p.res2 = res2; p.ln2_g = ln2_g; p.ln2_b = ln2_b;
    p.zero_from = nullptr; p.src = src; p.pad_id = pad_id; p.tile_list = tile_list; p.tile_count = tile_count;
    p.Y = Y; p.err = g_err_flag; p.trace = (g_trace && g_trace_count++ == g_trace_pick) ? g_trace : nullptr;
    const int n_tiles = B * ((T + TM - 1) / TM);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    if (!res2) return launch_mode<MODE_DWCONV, false, true, false>(p, grid, s);
    return gather_x ? launch_mode<MODE_DWCONV, true, true, true>(p, grid, s) : launch_mode<MODE_DWCONV, true, false, true>(p, grid, s);
}

// device flag raised by any tcgen05 kernel whose bounded mbarrier wait timed out (shared with es_umma_enc.cu)
int* umma_err_flag() {
    int*& flag = g_err_flags.get();
    if (!flag) {
        if (cudaMalloc(&flag, sizeof(int)) != cudaSuccess) { flag = nullptr; return nullptr; }
        cudaMemset(flag, 0, sizeof(int));
    }
    return flag;
}

// debug: subsequent tcgen05 decoder launches stamp clock64() per role/tile/event into buf (CTA 0)
void umma_dec_set_trace(long long* buf) {
    g_trace = buf;
    g_trace_count = 0;
    const char* e = getenv("ES_TRACE_LAUNCH");
    g_trace_pick = e ? atoi(e) : 0;
}

// Reads (and clears) the device-side mbarrier-timeout flag; synchronises the stream.
int umma_dec_check_errors(cudaStream_t s) {
    int* const g_err_flag = g_err_flags.get();
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

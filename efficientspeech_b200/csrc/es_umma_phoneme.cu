// The whole phoneme side of the acoustic path in ONE kernel (sm_100a: tcgen05 + TMEM + bulk copies):
// PhonemeEncoder.forward up to the feature upsampler (layers/networks.py:336-384) for the tiny
// geometry (d = 32: C = 32 / 64, heads 1 / 2, kernel 3, expansion 1) and N <= 128 phonemes.
//
// The per-layer path spends ~12 us per launch on 18 dependent launches (launch gap, prologue, 3-4 serial
// 64-row tiles per CTA, drain) for a few hundred KB of activations per utterance.  Here one CTA owns one
// utterance at a time (persistent loop, one CTA per SM) and keeps every activation on chip.
//
// 512 threads = 16 warps: phoneme row r <-> TMEM lane r is shared by FOUR threads (warps w, w+4, w+8, w+12
// all address lane quarter w % 4), each owning one quarter of the columns of whatever layer is being
// finished.  (A first version with one thread per row was correct but latency-bound: 4 warps per SM cannot
// hide a ~20 k-instruction dependent chain per row.)  Every dense layer is
//     the row's four threads write their K/4 slices of the A operand (split fp16 hi/lo, UMMA canonical
//     K-major row-panel layout: addr(row, k) = (k/8)*LBO + row*16 + (k%8)*2)
//     -> fence + __syncthreads -> one elected thread issues the tcgen05.mma's (M = 128, 3 per K step:
//     hi*hi + hi*lo + lo*hi, fp32 accumulate in TMEM) and commits to an mbarrier
//     -> every thread reads its N/4 accumulator columns (tcgen05.ld 32x32b) and runs the epilogue in registers;
// row statistics (LayerNorm, softmax max / sum, the scalar heads) are combined across the four threads through
// a ping-pong shared-memory array and one __syncthreads.  Conv taps are descriptor row shifts over a tile with
// one zero halo row on each side (es_umma_enc.cu); attention is S = Q K^T, softmax out of TMEM, O = P V
// (es_umma_attn.cu).  Weights are the packed split-fp16 images of the per-layer kernels, streamed through two
// 48 KB shared-memory buffers by bulk copies issued two layers ahead.  Level-1 rows (n1 = ceil(N/2) <= 64)
// live in rows 0..n1-1; the M = 128 GEMMs simply carry zero rows.  Where the row -> thread map changes, rows are
// handed over through a small per-utterance scratch array in global memory (L2; stride-2 merge conv -> residual of
// block 1) or through shared memory (the U rows of Fuse's stride-2 transposed-conv scatter).
//
// Layer list per utterance (reference lines in es_api.cu next to the per-layer launches):
//   embed+merge0 (3 table gathers) | qkv0, attention0, proj0+res+LN1+mask, conv3(ffn1)+GELU, ffn2+res+LN2+mask |
//   merge1 (stride 2) | qkv1 (q, k, v), attention1 x 2 heads, proj1+.., ffn1, ffn2 | Fuse (U, A0 + scatter) |
//   3 predictors x (conv3+ReLU+LN+ReLU, conv3+ReLU+scalar head [+LN2 for duration]) |
//   bucketize + embeddings + concat -> fused4, duration rounding, block scan -> dur_cum, mel_len.
#include <math.h>
#include <string.h>

#include <type_traits>

#include "es_common.cuh"
#include "es_kernels.cuh"
#include "es_umma.cuh"

namespace es {
namespace {

using namespace umma;

constexpr int PM = 128;                          // rows = TMEM lanes
constexpr int NTHR = 4 * PM;                     // four threads per row (column quarters)
constexpr int D0 = 32, D1 = 64;                  // channel widths of the two pyramid levels
constexpr uint32_t PANEL = PM * 16;              // 2048: one K panel (8 elements) of 128 rows, no halo
constexpr uint32_t XA_LBO = (PM + 2) * 16;       // 2080: one K panel of 130 rows (zero halo row on each side)
constexpr uint32_t XA_PLANE = 8 * XA_LBO;        // 16640: up to K = 64
constexpr uint32_t Y1_PLANE = 4 * XA_LBO;        // 8320: K = 32 halo tile
constexpr uint32_t Y1_TILE = 2 * Y1_PLANE;       // 16640
constexpr uint32_t WB_BYTES = 48 * 1024;

constexpr uint32_t OFF_OP = 0;                                 // 64 KB: attention Q/K then P | proj1 A tile | predictor y1 tiles
constexpr uint32_t OFF_VT = OFF_OP + 65536;                    // 16 KB: V^T hi, lo
constexpr uint32_t OFF_XA = OFF_VT + 16384;                    // halo A tile (hi plane, lo plane)
constexpr uint32_t OFF_WB = OFF_XA + 2 * XA_PLANE;             // two weight buffers
constexpr uint32_t OFF_RED = OFF_WB + 2 * WB_BYTES;            // row reductions: [2][4][128] float2
constexpr uint32_t OFF_MISC = OFF_RED + 2 * 4 * PM * 8;        // scan scratch
constexpr uint32_t OFF_BINS = OFF_MISC + 64;                   // pitch | energy bucket boundaries: 2 x 32 floats
constexpr uint32_t OFF_BAR = OFF_BINS + 2 * 32 * 4;            // bar_mma, bar_w[2], tmem slot
constexpr uint32_t PH_SMEM = OFF_BAR + 64;
static_assert(OFF_WB % 16 == 0 && OFF_XA % 16 == 0, "bulk copies need 16-byte aligned destinations");
static_assert(PH_SMEM <= 227 * 1024, "shared memory budget");
static_assert(3 * Y1_TILE <= 65536, "predictor tiles live in the attention region");

constexpr int NW = 15;                           // weight loads per utterance (see wload)

struct PhonemeParams {
    int B, N, n1, pool, n_symbols;
    const int32_t* ids;
    const uint8_t* mask;
    const float* pitch_tgt; const float* energy_tgt; const int32_t* dur_tgt;
    es_enc_block_w_t enc[2];
    const void* fuse_u_h16; const float* fuse_gb; const void* fuse_a0_h16; const float* fuse_c;
    es_predictor_w_t pred[3];                    // pitch, energy, duration
    float* pitch_pred; float* energy_pred; float* dur_pred; float* fused4;
    int32_t* dur_int; int32_t* dur_cum; int32_t* mel_len;
    float* sc_xm1;                               // [B][n1][64]  block-1 input rows (residual of proj1)
    float scale_log2e;                           // (C // H)^-0.5 * log2(e): 32^-0.5 at both levels
    int* err;
    long long* trace;                            // debug: clock64 stamps of CTA 0 / thread 0, [2 utterances][128 events], or null
};

__device__ __forceinline__ float ex2f_approx(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
    return e;
}

// Latency-critical waits (an MMA round trip is ~0.3 us, and ~25 of them are serial per utterance): plain try_wait
// spin instead of the suspend-hint form of es_umma.cuh, whose wake-up granularity showed up as ~1 us per wait.
// Out of line: the kernel is one long straight-line pass per utterance and instruction fetch is a measured cost.
// Bounded like every wait in this library.
__device__ __noinline__ bool spin_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 28); ++i) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

#define PHASE_SYNC()               \
    do {                           \
        fence_proxy_async_smem();  \
        tc_fence_before_sync();    \
        __syncthreads();           \
        tc_fence_after_sync();     \
    } while (0)

// NV (8 | 16 | 32) fp32 values = NV/8 consecutive K panels starting at panel pc0 -> split fp16 hi/lo rows of a
// row-panel tile (tile row `trow`)
template <int NV>
__device__ __forceinline__ void stage_cols(uint8_t* tile, uint32_t lbo, uint32_t plane, int trow, int pc0, const float* v, bool live) {
#pragma unroll
    for (int pc = 0; pc < NV / 8; ++pc) {
        float a[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = live ? v[8 * pc + e] : 0.f;
        uint4 hi, lo;
        split8(a, hi, lo);
        const uint32_t off = (uint32_t)(pc0 + pc) * lbo + (uint32_t)trow * 16u;
        *reinterpret_cast<uint4*>(tile + off) = hi;
        *reinterpret_cast<uint4*>(tile + plane + off) = lo;
    }
}

// NV consecutive accumulator columns of this thread's TMEM lane
template <int NV>
__device__ __forceinline__ void tm_load(uint32_t taddr, float* v) {
    static_assert(NV == 8 || NV == 16 || NV == 32, "8, 16 or 32 columns");
    uint32_t rr[NV];
    if constexpr (NV == 8) tmem_ld32x8(taddr, rr);
    else if constexpr (NV == 16) tmem_ld32x16(taddr, rr);
    else tmem_ld32(taddr, rr);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = __uint_as_float(rr[j]);
}

// D[tmem_d] = A . B^T, called by all lanes of warp 0 with warp-uniform arguments.  A: row-panel tile (K panel stride a_lbo,
// lo plane at +a_plane); conv tap t reads it shifted by t rows (16 bytes).  B: [N rows][K] in the same layout
// (panel stride b_lbo, lo plane at +b_plane, tap t at +t*b_tap).  The four descriptors advance by a constant per
// K step (start-address field, 16-byte units), so a K step costs 4 adds + 3 tcgen05.mma.  Inlined, with an
// elect.sync predicate around the MMAs only, so the descriptors live in UNIFORM registers: as an out-of-line
// function taking a per-thread `elected` flag every MMA paid a chain of R2UR moves (~100 cycles per MMA in the
// timeline).  The loops stay rolled to keep the code small (instruction fetch is a measured cost here).
template <int KSTEPS, int TAPS>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_addr, uint32_t a_lbo, uint32_t a_plane,
                                           uint32_t b_addr, uint32_t b_lbo, uint32_t b_plane, uint32_t b_tap, int N) {
    const bool elected = elect_one();
    const uint32_t idesc = make_idesc_f16(PM, N);
    const uint64_t dah0 = make_smem_desc(a_addr, a_lbo, 128u), dal0 = make_smem_desc(a_addr + a_plane, a_lbo, 128u);
    const uint64_t dbh0 = make_smem_desc(b_addr, b_lbo, 128u), dbl0 = make_smem_desc(b_addr + b_plane, b_lbo, 128u);
#pragma unroll
    for (int t = 0; t < TAPS; ++t) {
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            // offsets in the 16-byte units of the descriptors' start-address field
            const uint64_t da = (uint64_t)(((uint32_t)t * 16u + (uint32_t)(2 * ks) * a_lbo) >> 4);
            const uint64_t db = (uint64_t)(((uint32_t)t * b_tap + (uint32_t)(2 * ks) * b_lbo) >> 4);
            if (elected) {
                mma_f16_ss(tmem_d, dah0 + da, dbh0 + db, idesc, (t | ks) ? 1u : 0u);
                mma_f16_ss(tmem_d, dah0 + da, dbl0 + db, idesc, 1u);
                mma_f16_ss(tmem_d, dal0 + da, dbh0 + db, idesc, 1u);
            }
        }
    }
}

// Three independent layers of identical geometry (the three predictors): problem i uses A tile a_addr + i*a_step,
// weights w_addr + i*w_step, accumulator columns tmem_d + i*N.  Their MMAs are interleaved so that consecutive
// tcgen05.mma's never target the same accumulator: back-to-back MMAs into ONE accumulator serialise on the
// accumulate dependency (~117 cycles each at M = 128 whatever N is; tools/trace_phoneme.py), independent ones pipeline.
__device__ __forceinline__ void issue_layer_x3(uint32_t tmem_d, uint32_t a_addr, uint32_t a_step, uint32_t a_lbo,
                                               uint32_t a_plane, uint32_t w_addr, uint32_t w_step, int K, int N, int taps) {
    const bool elected = elect_one();
    const uint32_t idesc = make_idesc_f16(PM, N);
    const uint32_t b_lbo = (uint32_t)N * 16u, w_plane = (uint32_t)N * (uint32_t)K * 2u;
    uint32_t acc = 0;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const uint32_t ao = (uint32_t)t * 16u + (uint32_t)(2 * ks) * a_lbo;
            const uint32_t bo = (uint32_t)(2 * t) * w_plane + (uint32_t)(2 * ks) * b_lbo;
#pragma unroll
            for (int term = 0; term < 3; ++term) {               // hi*hi, hi*lo, lo*hi
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const uint64_t da = make_smem_desc(a_addr + (uint32_t)i * a_step + (term == 2 ? a_plane : 0u) + ao, a_lbo, 128u);
                    const uint64_t db = make_smem_desc(w_addr + (uint32_t)i * w_step + (term == 1 ? w_plane : 0u) + bo, b_lbo, 128u);
                    if (elected) mma_f16_ss(tmem_d + (uint32_t)(i * N), da, db, idesc, term == 0 ? acc : 1u);
                }
            }
            acc = 1;
        }
    }
}
// dense layer with packed weights [taps][hi, lo][K/8][N][8]
template <int K, int TAPS>
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_addr, uint32_t a_lbo, uint32_t a_plane, uint32_t w_addr, int N) {
    const uint32_t w_plane = (uint32_t)N * (uint32_t)K * 2u;
    issue_gemm<K / 16, TAPS>(tmem_d, a_addr, a_lbo, a_plane, w_addr, (uint32_t)N * 16u, w_plane, 2u * w_plane, N);
}

// NV (multiple of 4) consecutive per-channel parameters (bias, LayerNorm gain, ...) with 128-bit loads; the
// packed parameter vectors are 256-byte aligned and every thread's first column is a multiple of 8
template <int NV>
__device__ __forceinline__ void ldv(const float* __restrict__ src, float* v) {
#pragma unroll
    for (int c4 = 0; c4 < NV / 4; ++c4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src) + c4);
        v[4 * c4] = t.x; v[4 * c4 + 1] = t.y; v[4 * c4 + 2] = t.z; v[4 * c4 + 3] = t.w;
    }
}

__device__ __forceinline__ int bucket_left(const float* bins, int nb, float v) {      // bins: shared memory
    int lo = 0, hi = nb;                          // #{j : bins[j] < v}   (torch.bucketize, right=False; networks.py:130-141)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (bins[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// TRACE: compiled-in clock64 stamps (tools/trace_phoneme.py); the production instantiation carries none
template <bool TRACE>
__global__ void __launch_bounds__(NTHR, 1)
umma_phoneme_kernel(const __grid_constant__ PhonemeParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rq = warp & 3, cq = warp >> 2;                   // TMEM lane quarter, column quarter
    const int r = 32 * rq + lane;                              // this thread's row
    uint8_t* op = smem + OFF_OP;
    uint8_t* vt = smem + OFF_VT;
    uint8_t* xa = smem + OFF_XA;
    float2* red = reinterpret_cast<float2*>(smem + OFF_RED);
    int* s_wtot = reinterpret_cast<int*>(smem + OFF_MISC);
    float* s_bins = reinterpret_cast<float*>(smem + OFF_BINS);
    float* u_s = reinterpret_cast<float*>(smem + OFF_OP);      // Fuse: U rows [n1][96] (the attention region is free by then)
    const uint32_t bar_mma = smem_u32(smem + OFF_BAR);
    const uint32_t bar_w = bar_mma + 8;                        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 32);
    const uint32_t xa_addr = smem_u32(xa), op_addr = smem_u32(op), vt_addr = smem_u32(vt);
    const uint32_t wb_addr = smem_u32(smem + OFF_WB);

    const int N = p.N, n1 = p.n1;
    const int my_utts = (p.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_w = my_utts * NW;

    // weight load g of this CTA (phase g % NW of utterance g / NW) into buffer g & 1; thread 0 only
    auto wload = [&](int g) {
        if (g >= total_w) return;
        const int ph = g % NW;
        const uint32_t dst = wb_addr + (uint32_t)(g & 1) * WB_BYTES;
        const uint32_t bar = bar_w + 8u * (uint32_t)(g & 1);
        auto one = [&](const void* src, uint32_t bytes) {
            mbar_arrive_expect_tx(bar, bytes);
            bulk_g2s(dst, src, bytes, bar);
        };
        switch (ph) {
            case 0: one(p.enc[0].qkv_w_h16, 96u * D0 * 4u); break;
            case 1: one(p.enc[0].proj_w_h16, D0 * D0 * 4u); break;
            case 2: one(p.enc[0].ffn1_w_h16, 3u * D0 * D0 * 4u); break;
            case 3: one(p.enc[0].ffn2_w_h16, D0 * D0 * 4u); break;
            case 4: one(p.enc[1].merge_w_h16, D1 * D0 * 4u); break;
            case 5: case 6: case 7: {
                // q | k | v slice (128 of the 384 output rows) of the [2][8][384][8] image -> compact [2][8][128][8]
                const uint8_t* img = reinterpret_cast<const uint8_t*>(p.enc[1].qkv_w_h16) + (uint32_t)(ph - 5) * 2048u;
                mbar_arrive_expect_tx(bar, 32768u);
                for (int pl = 0; pl < 2; ++pl)
                    for (int pc = 0; pc < 8; ++pc)
                        bulk_g2s(dst + (uint32_t)pl * 16384u + (uint32_t)pc * 2048u, img + (size_t)pl * 49152u + (size_t)pc * 6144u, 2048u, bar);
                break;
            }
            case 8: one(p.enc[1].proj_w_h16, D1 * 128u * 4u); break;
            case 9: one(p.enc[1].ffn1_w_h16, 3u * D1 * D1 * 4u); break;
            case 10: one(p.enc[1].ffn2_w_h16, D1 * D1 * 4u); break;
            case 11: one(p.fuse_u_h16, 96u * D1 * 4u); break;
            case 12: one(p.fuse_a0_h16, D0 * D0 * 4u); break;
            case 13:
                mbar_arrive_expect_tx(bar, 3u * 12288u);
                for (int i = 0; i < 3; ++i) bulk_g2s(dst + (uint32_t)i * 12288u, p.pred[i].conv1_w_h16, 12288u, bar);
                break;
            default:
                mbar_arrive_expect_tx(bar, 3u * 12288u);
                for (int i = 0; i < 3; ++i) bulk_g2s(dst + (uint32_t)i * 12288u, p.pred[i].conv2_w_h16, 12288u, bar);
                break;
        }
    };

    // ---- one-time setup -------------------------------------------------------------------------
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);       // [0,128) S / layer accumulators, [128,512) q | k | v (O over q)
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_w, 1);
        mbar_init(bar_w + 8, 1);
        fence_mbar_init();
    }
    if (tid < 64) s_bins[tid] = (tid & 31) < D0 - 1 ? __ldg((tid < 32 ? p.pred[0].bins : p.pred[1].bins) + (tid & 31)) : 0.f;
    // zero halo rows (tile rows 0 and 129) of the A tile: never written afterwards
    if (tid < 32) {
        const int pc = tid & 7, pl = (tid >> 3) & 1, which = tid >> 4;
        *reinterpret_cast<uint4*>(xa + (uint32_t)pl * XA_PLANE + (uint32_t)pc * XA_LBO + (which ? (PM + 1) * 16u : 0u)) = make_uint4(0u, 0u, 0u, 0u);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(rq * 32) << 16);
    if (tid == 0) { wload(0); wload(1); }                      // weights do not depend on the predecessor kernel
    pdl_launch_dependents();
    pdl_wait();

    bool failed = false;
    uint32_t mph = 0;                                           // bar_mma phase parity
    int wg = 0;                                                 // next weight phase to be consumed (global index)
    int flip = 0;                                               // ping-pong half of `red`
    // debug timeline (tools/trace_phoneme.py): thread 0 of CTA 0 stamps every GEMM completion (odd codes) and phase sync
    int tr_k = 0, tr_u = 0;
    auto stamp = [&](int code) {
        if (TRACE && p.trace && blockIdx.x == 0 && tid == 0 && tr_u < 2 && tr_k < 128) {
            p.trace[tr_u * 256 + 2 * tr_k] = clock64();
            p.trace[tr_u * 256 + 2 * tr_k + 1] = code;
            ++tr_k;
        }
    };

    // warp 0 waits for weight load `wg`; returns its shared-memory address
    auto w_wait = [&]() -> uint32_t {
        stamp(4);
        if (warp == 0 && !spin_wait(bar_w + 8u * (uint32_t)(wg & 1), (uint32_t)(wg >> 1) & 1u)) failed = true;
        stamp(5);
        return wb_addr + (uint32_t)(wg & 1) * WB_BYTES;
    };
    // commit the issued MMAs, wait for them, then refill the weight buffer they used (two phases ahead)
    auto gemm_done = [&](bool used_weights) {
        if (warp == 0) {
            if (elect_one()) mma_commit(bar_mma);
            __syncwarp();
        }
        if (!spin_wait(bar_mma, mph)) failed = true;
        mph ^= 1;
        tc_fence_after_sync();
        if (used_weights) {
            if (tid == 0) wload(wg + 2);
            ++wg;
        }
        stamp(used_weights ? 1 : 3);
    };
    // combine two per-thread partials over the four threads of a row (same order in all four -> identical results).
    // One barrier per call, among the four warps of this lane quarter only (rows of other quarters are independent):
    // the buffer halves alternate, and a half is rewritten only after the barrier of the following call, which
    // every reader of this call has passed.
    auto row_sum2 = [&](float a, float b2) -> float2 {
        float2* bf = red + flip * (4 * PM);
        flip ^= 1;
        bf[cq * PM + r] = make_float2(a, b2);
        named_bar_sync(1 + rq, PM);                          // only the four warps that share these 32 rows
        stamp(6);
        const float2 t0 = bf[r], t1 = bf[PM + r], t2 = bf[2 * PM + r], t3 = bf[3 * PM + r];
        return make_float2((t0.x + t1.x) + (t2.x + t3.x), (t0.y + t1.y) + (t2.y + t3.y));
    };
    auto row_max = [&](float a) -> float {
        float2* bf = red + flip * (4 * PM);
        flip ^= 1;
        bf[cq * PM + r].x = a;
        named_bar_sync(1 + rq, PM);
        return fmaxf(fmaxf(bf[r].x, bf[PM + r].x), fmaxf(bf[2 * PM + r].x, bf[3 * PM + r].x));
    };
    // LayerNorm over n_tot = 4 * NV columns of the row; this thread holds NV of them, g / be point at its first column
    // (single-pass statistics like the per-layer tensor-core kernels: inputs are O(1))
    auto ln_cols = [&](auto nv_tag, float* v, const float* g, const float* be) {
        constexpr int NV = decltype(nv_tag)::value;
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) { s += v[j]; q = fmaf(v[j], v[j], q); }
        const float2 t = row_sum2(s, q);
        const float inv_n = 1.f / (4 * NV);
        const float mean = t.x * inv_n;
        const float rstd = rsqrtf(fmaxf(fmaf(t.y, inv_n, -mean * mean), 0.f) + kLnEps);
        float gg[NV], bb[NV];
        ldv<NV>(g, gg);
        ldv<NV>(be, bb);
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = fmaf((v[j] - mean) * rstd, gg[j], bb[j]);
    };
    using I8 = std::integral_constant<int, 8>;
    using I16 = std::integral_constant<int, 16>;

    // softmax(Q K^T) V for one head: q/k/v rows are in TMEM columns qc/kc/vc (C wide), output -> columns oc.
    // This thread owns CV = C/4 channels of its row and SV = NK/4 keys of its score row.
    auto attention = [&](auto c_tag, auto nk_tag, const int n, const int qc, const int kc, const int vc, const int oc) {
        constexpr int C = decltype(c_tag)::value, NK = decltype(nk_tag)::value;
        constexpr int CV = C / 4, SV = NK / 4;
        constexpr uint32_t qk_plane = (uint32_t)(C >> 3) * PANEL;
        constexpr uint32_t vt_lbo = (uint32_t)C * 16u, vt_plane = (uint32_t)(NK >> 3) * vt_lbo;
        constexpr uint32_t p_plane = (uint32_t)(NK >> 3) * PANEL;
        // 1. stage Q, K (row panels) and V^T; rows >= n are zero rows of the GEMM that produced them
        {
            float q[CV], k[CV], v[CV];
            tm_load<CV>(trow + (uint32_t)(qc + CV * cq), q);
            tm_load<CV>(trow + (uint32_t)(kc + CV * cq), k);
            tm_load<CV>(trow + (uint32_t)(vc + CV * cq), v);
            stage_cols<CV>(op, PANEL, qk_plane, r, (CV / 8) * cq, q, true);
            stage_cols<CV>(op + 2 * qk_plane, PANEL, qk_plane, r, (CV / 8) * cq, k, true);
            if (r < NK) {
#pragma unroll
                for (int e = 0; e < CV; ++e) {
                    const __half vh = __float2half_rn(v[e]);
                    const __half vl = __float2half_rn(v[e] - __half2float(vh));
                    const uint32_t off = (uint32_t)(r >> 3) * vt_lbo + (uint32_t)(CV * cq + e) * 16u + (uint32_t)(r & 7) * 2u;
                    *reinterpret_cast<__half*>(vt + off) = vh;
                    *reinterpret_cast<__half*>(vt + vt_plane + off) = vl;
                }
            }
        }
        PHASE_SYNC(); stamp(2);
        // 2. S = Q K^T  -> columns [0, NK)
        if (warp == 0) issue_gemm<C / 16, 1>(tmem, op_addr, PANEL, qk_plane, op_addr + 2 * qk_plane, PANEL, qk_plane, 0u, NK);
        gemm_done(false);
        // 3. softmax over the n keys (unmasked: blocks.py:59-63), P -> split fp16 over the dead Q/K tiles
        {
            const float sc = p.scale_log2e;
            float s[SV];
            tm_load<SV>(trow + (uint32_t)(SV * cq), s);
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < SV; ++j)
                if (SV * cq + j < n) mx = fmaxf(mx, s[j]);
            mx = row_max(mx);
            const float nm = -mx * sc;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < SV; ++j) {
                const float e = ex2f_approx(fmaf(s[j], sc, nm));
                s[j] = (SV * cq + j < n) ? e : 0.f;
                sum += s[j];
            }
            const float inv = 1.f / row_sum2(sum, 0.f).x;
#pragma unroll
            for (int j = 0; j < SV; ++j) s[j] *= inv;
            stage_cols<SV>(op, PANEL, p_plane, r, (SV / 8) * cq, s, true);
        }
        PHASE_SYNC(); stamp(2);
        // 4. O = P V  -> columns [oc, oc + C)
        if (warp == 0) issue_gemm<NK / 16, 1>(tmem + (uint32_t)oc, op_addr, PANEL, p_plane, vt_addr, vt_lbo, vt_plane, 0u, C);
        gemm_done(false);
    };

    const int c8 = 8 * cq, c16 = 16 * cq;                       // this thread's first column of a 32- / 64-wide row

    // x0 = sum_tau Tab[tau][id[t + tau - 1]]   (embedding + merge conv + 1x1 folded into 3 gather tables);
    // two dependent global loads, so the row of the NEXT utterance is fetched while this one runs its predictors
    auto gather_x0 = [&](int b, float* x0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) x0[c] = 0.f;
        if (r < N) {
            for (int tau = 0; tau < 3; ++tau) {
                const int ti = r + tau - 1;
                if (ti < 0 || ti >= N) continue;
                int id = __ldg(p.ids + (size_t)b * N + ti);
                id = min(max(id, 0), p.n_symbols - 1);
                const float4* tp = reinterpret_cast<const float4*>(p.enc[0].merge_w + ((size_t)tau * p.n_symbols + id) * D0 + c8);
                const float4 v0 = __ldg(tp), v1 = __ldg(tp + 1);
                x0[0] += v0.x; x0[1] += v0.y; x0[2] += v0.z; x0[3] += v0.w;
                x0[4] += v1.x; x0[5] += v1.y; x0[6] += v1.z; x0[7] += v1.w;
            }
        }
    };
    float x0n[8];
    if (my_utts > 0) gather_x0((int)blockIdx.x, x0n);

    for (int ui = 0; ui < my_utts; ++ui) {
        const int b = (int)blockIdx.x + ui * (int)gridDim.x;
        const size_t row = (size_t)b * N + r;
        const bool act0 = r < N, act1 = r < n1;
        tr_k = 0; tr_u = ui;
        stamp(0);
        const bool pad0 = act0 && p.mask && p.mask[row];
        bool pad1 = false;                                      // pooled mask of block 1 (blocks.py:51-57)
        if (act1 && p.mask) {
            for (int q = 0; q < p.pool; ++q) {
                const int t = r * p.pool + q;
                pad1 = pad1 || (t < N ? p.mask[(size_t)b * N + t] != 0 : true);
            }
        }

        // ================================================================= block 0 (8 columns per thread)
        float x0[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) x0[c] = x0n[c];
        // per-row scalars of the variance stage: requested now, consumed ~70 us later
        float tgt_p = 0.f, tgt_e = 0.f;
        int tgt_d = 0;
        if (act0) {
            if (p.pitch_tgt) tgt_p = __ldg(p.pitch_tgt + row);
            if (p.energy_tgt) tgt_e = __ldg(p.energy_tgt + row);
            if (p.dur_tgt) tgt_d = __ldg(p.dur_tgt + row);
        }
        stage_cols<8>(xa, XA_LBO, XA_PLANE, r + 1, cq, x0, act0);
        PHASE_SYNC(); stamp(2);
        {   // qkv0: [32] -> [96] into columns [128, 224)
            const uint32_t w = w_wait();
            if (warp == 0) issue_layer<D0, 1>(tmem + 128, xa_addr + 16, XA_LBO, XA_PLANE, w, 96);
            gemm_done(true);
        }
        attention(std::integral_constant<int, D0>{}, std::integral_constant<int, PM>{}, N, 128, 160, 192, 128);
        float x1[8];
        {   // proj0 + residual + LN1 + mask
            float o[8];
            tm_load<8>(trow + (uint32_t)(128 + c8), o);
            stage_cols<8>(xa, XA_LBO, XA_PLANE, r + 1, cq, o, act0);
            PHASE_SYNC(); stamp(2);
            const uint32_t w = w_wait();
            if (warp == 0) issue_layer<D0, 1>(tmem, xa_addr + 16, XA_LBO, XA_PLANE, w, D0);
            gemm_done(true);
            tm_load<8>(trow + (uint32_t)c8, x1);
            float pb[8];
            ldv<8>(p.enc[0].proj_b + c8, pb);
#pragma unroll
            for (int c = 0; c < 8; ++c) x1[c] += pb[c] + x0[c];
            ln_cols(I8{}, x1, p.enc[0].ln1_g + c8, p.enc[0].ln1_b + c8);
            if (pad0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) x1[c] = 0.f;
            }
            stage_cols<8>(xa, XA_LBO, XA_PLANE, r + 1, cq, x1, act0);
            PHASE_SYNC(); stamp(2);
        }
        float f0[8];
        {   // MixFFN: conv3 (mlp1 folded) + GELU, mlp2 + residual + LN2 + mask
            uint32_t w = w_wait();
            if (warp == 0) issue_layer<D0, 3>(tmem, xa_addr, XA_LBO, XA_PLANE, w, D0);
            gemm_done(true);
            float h[8];
            tm_load<8>(trow + (uint32_t)c8, h);
            const float m0 = (r - 1 >= 0 && r - 1 < N) ? 1.f : 0.f, m2 = (r + 1 < N) ? 1.f : 0.f;
            float fb[8], t0[8], t1[8], t2[8];
            ldv<8>(p.enc[0].ffn1_b + c8, fb);
            ldv<8>(p.enc[0].ffn1_tapb + c8, t0);
            ldv<8>(p.enc[0].ffn1_tapb + D0 + c8, t1);
            ldv<8>(p.enc[0].ffn1_tapb + 2 * D0 + c8, t2);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float v = h[c] + fb[c];
                v = fmaf(m0, t0[c], v);
                v += t1[c];
                v = fmaf(m2, t2[c], v);
                h[c] = gelu_erf_f(v);
            }
            stage_cols<8>(xa, XA_LBO, XA_PLANE, r + 1, cq, h, act0);
            PHASE_SYNC(); stamp(2);
            w = w_wait();
            if (warp == 0) issue_layer<D0, 1>(tmem, xa_addr + 16, XA_LBO, XA_PLANE, w, D0);
            gemm_done(true);
            tm_load<8>(trow + (uint32_t)c8, f0);
            float f2b[8];
            ldv<8>(p.enc[0].ffn2_b + c8, f2b);
#pragma unroll
            for (int c = 0; c < 8; ++c) f0[c] += f2b[c] + x1[c];
            ln_cols(I8{}, f0, p.enc[0].ln2_g + c8, p.enc[0].ln2_b + c8);
            if (pad0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) f0[c] = 0.f;
            }
            stage_cols<8>(xa, XA_LBO, XA_PLANE, r + 1, cq, f0, act0);
            PHASE_SYNC(); stamp(2);
        }
        // ================================================================= block 1 (16 columns per thread)
        {   // merge conv (1 tap, stride 2): xm1[j] = W feat0[2j]; computed for every row, even rows are kept
            const uint32_t w = w_wait();
            if (warp == 0) issue_layer<D0, 1>(tmem, xa_addr + 16, XA_LBO, XA_PLANE, w, D1);
            gemm_done(true);
            float y[16];
            tm_load<16>(trow + (uint32_t)c16, y);
            if (act0 && !(r & 1)) {
                const int j = r >> 1;                            // < n1
                float4* dst = reinterpret_cast<float4*>(p.sc_xm1 + ((size_t)b * n1 + j) * D1 + c16);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) dst[c4] = make_float4(y[4 * c4], y[4 * c4 + 1], y[4 * c4 + 2], y[4 * c4 + 3]);
                stage_cols<16>(xa, XA_LBO, XA_PLANE, j + 1, 2 * cq, y, true);
            }
            if (r >= n1) stage_cols<16>(xa, XA_LBO, XA_PLANE, r + 1, 2 * cq, y, false);      // zero rows n1..127
            PHASE_SYNC(); stamp(2);
        }
        for (int s = 0; s < 3; ++s) {   // q | k | v of both heads: [64] -> [128] into columns 128 + 128 s
            const uint32_t w = w_wait();
            if (warp == 0) issue_layer<D1, 1>(tmem + 128 + 128 * s, xa_addr + 16, XA_LBO, XA_PLANE, w, 128);
            gemm_done(true);
        }
        for (int hd = 0; hd < 2; ++hd)
            attention(std::integral_constant<int, D1>{}, std::integral_constant<int, 64>{}, n1,
                      128 + 64 * hd, 256 + 64 * hd, 384 + 64 * hd, 128 + 64 * hd);
        float x1b[16];
        {   // proj1: A = [O_0 | O_1] (K = 128) in the attention region; + residual + LN1 + mask
            float o[32];
            tm_load<32>(trow + (uint32_t)(128 + 32 * cq), o);
            stage_cols<32>(op, PANEL, 16 * PANEL, r, 4 * cq, o, act1);
            PHASE_SYNC(); stamp(2);
            const uint32_t w = w_wait();
            if (warp == 0) issue_layer<128, 1>(tmem, op_addr, PANEL, 16 * PANEL, w, D1);
            gemm_done(true);
            tm_load<16>(trow + (uint32_t)c16, x1b);
            if (act1) {
                const float4* rp = reinterpret_cast<const float4*>(p.sc_xm1 + ((size_t)b * n1 + r) * D1 + c16);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 v = __ldcg(rp + c4);
                    x1b[4 * c4] += v.x; x1b[4 * c4 + 1] += v.y; x1b[4 * c4 + 2] += v.z; x1b[4 * c4 + 3] += v.w;
                }
            }
            float pb[16];
            ldv<16>(p.enc[1].proj_b + c16, pb);
#pragma unroll
            for (int c = 0; c < 16; ++c) x1b[c] += pb[c];
            ln_cols(I16{}, x1b, p.enc[1].ln1_g + c16, p.enc[1].ln1_b + c16);
            if (pad1) {
#pragma unroll
                for (int c = 0; c < 16; ++c) x1b[c] = 0.f;
            }
            stage_cols<16>(xa, XA_LBO, XA_PLANE, r + 1, 2 * cq, x1b, act1);
            PHASE_SYNC(); stamp(2);
        }
        {   // MixFFN of block 1
            uint32_t w = w_wait();
            if (warp == 0) issue_layer<D1, 3>(tmem, xa_addr, XA_LBO, XA_PLANE, w, D1);
            gemm_done(true);
            float h[16];
            tm_load<16>(trow + (uint32_t)c16, h);
            const float m0 = (r - 1 >= 0 && r - 1 < n1) ? 1.f : 0.f, m2 = (r + 1 < n1) ? 1.f : 0.f;
            float fb[16], t0[16], t1[16], t2[16];
            ldv<16>(p.enc[1].ffn1_b + c16, fb);
            ldv<16>(p.enc[1].ffn1_tapb + c16, t0);
            ldv<16>(p.enc[1].ffn1_tapb + D1 + c16, t1);
            ldv<16>(p.enc[1].ffn1_tapb + 2 * D1 + c16, t2);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                float v = h[c] + fb[c];
                v = fmaf(m0, t0[c], v);
                v += t1[c];
                v = fmaf(m2, t2[c], v);
                h[c] = gelu_erf_f(v);
            }
            stage_cols<16>(xa, XA_LBO, XA_PLANE, r + 1, 2 * cq, h, act1);
            PHASE_SYNC(); stamp(2);
            w = w_wait();
            if (warp == 0) issue_layer<D1, 1>(tmem, xa_addr + 16, XA_LBO, XA_PLANE, w, D1);
            gemm_done(true);
            tm_load<16>(trow + (uint32_t)c16, h);
            float f2b[16];
            ldv<16>(p.enc[1].ffn2_b + c16, f2b);
#pragma unroll
            for (int c = 0; c < 16; ++c) h[c] += f2b[c] + x1b[c];
            ln_cols(I16{}, h, p.enc[1].ln2_g + c16, p.enc[1].ln2_b + c16);        // feat1
            if (pad1) {
#pragma unroll
                for (int c = 0; c < 16; ++c) h[c] = 0.f;
            }
            stage_cols<16>(xa, XA_LBO, XA_PLANE, r + 1, 2 * cq, h, act1);
            PHASE_SYNC(); stamp(2);
        }
        // ================================================================= Fuse (networks.py:189-219, folded)
        float fz[8];
        {
            uint32_t w = w_wait();                               // U = feat1 [G_0 | G_1 | G_2] + [g_0 | g_1 | g_2]
            if (warp == 0) issue_layer<D1, 1>(tmem, xa_addr + 16, XA_LBO, XA_PLANE, w, 96);
            gemm_done(true);
#pragma unroll 1
            for (int i = 0; i < 3; ++i) {   // 3 of the 12 8-column groups of U per thread
                const int c0 = 8 * (cq + 4 * i);
                float u[8];
                tm_load<8>(trow + (uint32_t)c0, u);
                if (act1) {
                    float4* dst = reinterpret_cast<float4*>(u_s + r * 96 + c0);
                    float gb[8];
                    ldv<8>(p.fuse_gb + c0, gb);
                    dst[0] = make_float4(u[0] + gb[0], u[1] + gb[1], u[2] + gb[2], u[3] + gb[3]);
                    dst[1] = make_float4(u[4] + gb[4], u[5] + gb[5], u[6] + gb[6], u[7] + gb[7]);
                }
            }
            stage_cols<8>(xa, XA_LBO, XA_PLANE, r + 1, cq, f0, act0);
            // rows of the K = 64 tile beyond panel 3 are not read by the K = 32 GEMMs that follow
            PHASE_SYNC(); stamp(2);                                        // also publishes the U rows (shared memory) to the CTA
            w = w_wait();                                        // fused = mask(c + A0 feat0 + stride-2 scatter of U)
            if (warp == 0) issue_layer<D0, 1>(tmem, xa_addr + 16, XA_LBO, XA_PLANE, w, D0);
            gemm_done(true);
            tm_load<8>(trow + (uint32_t)c8, fz);
            float fc[8];
            ldv<8>(p.fuse_c + c8, fc);
#pragma unroll
            for (int c = 0; c < 8; ++c) fz[c] += fc[c];
            if (act0) {
                for (int tau = r & 1; tau < 3; tau += 2) {
                    const int j = (r - tau) >> 1;
                    if (r < tau || j >= n1) continue;
                    const float4* up = reinterpret_cast<const float4*>(u_s + j * 96 + tau * D0 + c8);
                    const float4 v0 = up[0], v1 = up[1];
                    fz[0] += v0.x; fz[1] += v0.y; fz[2] += v0.z; fz[3] += v0.w;
                    fz[4] += v1.x; fz[5] += v1.y; fz[6] += v1.z; fz[7] += v1.w;
                }
            }
            if (pad0) {
#pragma unroll
                for (int c = 0; c < 8; ++c) fz[c] = 0.f;
            }
            if (act0) {
                float4* dst = reinterpret_cast<float4*>(p.fused4 + row * 128 + c8);
                dst[0] = make_float4(fz[0], fz[1], fz[2], fz[3]);
                dst[1] = make_float4(fz[4], fz[5], fz[6], fz[7]);
            }
            stage_cols<8>(xa, XA_LBO, XA_PLANE, r + 1, cq, fz, act0);
            PHASE_SYNC(); stamp(2);                                        // every U read is done: the region becomes the predictor tiles
        }
        // ================================================================= predictors (networks.py:151-165)
        float pred[3];
        {
            uint32_t w = w_wait();                               // conv1 of the three predictors -> columns 0 | 32 | 64
            if (warp == 0) issue_layer_x3(tmem, xa_addr, 0u, XA_LBO, XA_PLANE, w, 12288u, D0, D0, 3);
            gemm_done(true);
            // next utterance's embedding row: in flight during the predictor and variance stages
            if (ui + 1 < my_utts) gather_x0(b + (int)gridDim.x, x0n);
            // halo rows of the three predictor tiles
            if (tid < 48) {
                const int i = tid >> 4, pc = tid & 3, pl = (tid >> 2) & 1, which = (tid >> 3) & 1;
                *reinterpret_cast<uint4*>(op + (uint32_t)i * Y1_TILE + (uint32_t)pl * Y1_PLANE + (uint32_t)pc * XA_LBO +
                                          (which ? (PM + 1) * 16u : 0u)) = make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll 1
            for (int i = 0; i < 3; ++i) {
                float y[8];
                tm_load<8>(trow + (uint32_t)(32 * i + c8), y);
                float cb[8];
                ldv<8>(p.pred[i].conv1_b + c8, cb);
#pragma unroll
                for (int c = 0; c < 8; ++c) y[c] = fmaxf(y[c] + cb[c], 0.f);
                ln_cols(I8{}, y, p.pred[i].ln1_g + c8, p.pred[i].ln1_b + c8);
#pragma unroll
                for (int c = 0; c < 8; ++c) y[c] = fmaxf(y[c], 0.f);
                stage_cols<8>(op + (uint32_t)i * Y1_TILE, XA_LBO, Y1_PLANE, r + 1, cq, y, act0);
            }
            PHASE_SYNC(); stamp(2);
            w = w_wait();                                        // conv2 + ReLU, scalar head on the pre-LN2 values
            if (warp == 0) issue_layer_x3(tmem, op_addr, Y1_TILE, XA_LBO, Y1_PLANE, w, 12288u, D0, D0, 3);
            gemm_done(true);
            float y2[3][8], dot[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                tm_load<8>(trow + (uint32_t)(32 * i + c8), y2[i]);
                dot[i] = 0.f;
                float cb[8], lw[8];
                ldv<8>(p.pred[i].conv2_b + c8, cb);
                ldv<8>(p.pred[i].lin_w + c8, lw);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    y2[i][c] = fmaxf(y2[i][c] + cb[c], 0.f);
                    dot[i] = fmaf(y2[i][c], lw[c], dot[i]);
                }
            }
            const float2 d01 = row_sum2(dot[0], dot[1]);
            const float2 d2 = row_sum2(dot[2], 0.f);
            pred[0] = d01.x + __ldg(p.pred[0].lin_b);
            pred[1] = d01.y + __ldg(p.pred[1].lin_b);
            pred[2] = fmaxf(d2.x + __ldg(p.pred[2].lin_b), 0.f);  // duration: extra ReLU (networks.py:161-163)
            ln_cols(I8{}, y2[2], p.pred[2].ln2_g + c8, p.pred[2].ln2_b + c8);   // features = LN2(y), consumed for duration only
            if (act0) {
                float4* dst = reinterpret_cast<float4*>(p.fused4 + row * 128 + 3 * D0 + c8);
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                dst[0] = pad0 ? z : make_float4(y2[2][0], y2[2][1], y2[2][2], y2[2][3]);
                dst[1] = pad0 ? z : make_float4(y2[2][4], y2[2][5], y2[2][6], y2[2][7]);
                if (cq == 0) {
                    p.pitch_pred[row] = pred[0];
                    p.energy_pred[row] = pred[1];
                    p.dur_pred[row] = pred[2];
                }
            }
        }
        // ================================================================= variance embeddings, durations, scan
        {
            int dv = 0;
            if (act0) {
                float df = p.dur_tgt ? (float)tgt_d : rintf(pred[2]);                // torch.round: half to even (networks.py:379)
                if (pad0) df = 0.f;
                df = fminf(fmaxf(df, 0.f), 65535.f);
                dv = (int)df;
                if (cq == 0) p.dur_int[row] = dv;
                const float pv = p.pitch_tgt ? tgt_p : pred[0];
                const float ev = p.energy_tgt ? tgt_e : pred[1];
                const int pi = bucket_left(s_bins, D0 - 1, pv);
                const int ei = bucket_left(s_bins + 32, D0 - 1, ev);
                const float4* pt = reinterpret_cast<const float4*>(p.pred[0].table + (size_t)pi * D0 + c8);
                const float4* et = reinterpret_cast<const float4*>(p.pred[1].table + (size_t)ei * D0 + c8);
                float4* dp = reinterpret_cast<float4*>(p.fused4 + row * 128 + D0 + c8);
                float4* de = reinterpret_cast<float4*>(p.fused4 + row * 128 + 2 * D0 + c8);
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                dp[0] = pad0 ? z : __ldg(pt); dp[1] = pad0 ? z : __ldg(pt + 1);
                de[0] = pad0 ? z : __ldg(et); de[1] = pad0 ? z : __ldg(et + 1);
            }
            // block scan over the 128 rows: done by the cq == 0 threads (warps 0..3), warp-shuffle scan + 4 warp totals
            int incl = dv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (cq == 0 && lane == 31) s_wtot[rq] = incl;
            __syncthreads();
            if (cq == 0) {
                int prefix = 0;
                for (int w = 0; w < rq; ++w) prefix += s_wtot[w];
                if (act0) p.dur_cum[row] = prefix + incl;
                if (r == PM - 1) p.mel_len[b] = prefix + incl;
            }
            // s_wtot is rewritten only after several __syncthreads of the next utterance
        }
        stamp(9);
    }

    if (failed) atomicExch(p.err, 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

static long long* g_ph_trace = nullptr;
void umma_phoneme_set_trace(long long* buf) { g_ph_trace = buf; }

bool umma_phoneme_supported(const es_config_t& cfg, const es_weights_t& w, int N) {
    if (cfg.dim != D0 || cfg.head != 1 || cfg.kernel_size != 3 || cfg.expansion != 1) return false;
    if (N < 2 || N > PM) return false;
    if (!w.enc[0].qkv_w_h16 || !w.enc[0].proj_w_h16 || !w.enc[0].ffn1_w_h16 || !w.enc[0].ffn2_w_h16) return false;
    if (!w.enc[1].merge_w_h16 || !w.enc[1].qkv_w_h16 || !w.enc[1].proj_w_h16 || !w.enc[1].ffn1_w_h16 || !w.enc[1].ffn2_w_h16) return false;
    if (!w.fuse_u_h16 || !w.fuse_a0_h16) return false;
    const es_predictor_w_t* pr[3] = {&w.pitch, &w.energy, &w.duration};
    for (int i = 0; i < 3; ++i) if (!pr[i]->conv1_w_h16 || !pr[i]->conv2_w_h16) return false;
    return w.pitch.bins && w.pitch.table && w.energy.bins && w.energy.table;
}

// sc_xm1: B*n1*64 floats of scratch.  -1: outside the kernel's envelope.
int launch_umma_phoneme(const es_config_t& cfg, const es_weights_t& w, int B, int N, int n1, int pool,
                        const int32_t* ids, const uint8_t* mask, const float* pitch_tgt, const float* energy_tgt,
                        const int32_t* dur_tgt, float* pitch_pred, float* energy_pred, float* dur_pred, float* fused4,
                        int32_t* dur_int, int32_t* dur_cum, int32_t* mel_len, float* sc_xm1, cudaStream_t s) {
    if (!umma_phoneme_supported(cfg, w, N) || n1 > 64 || n1 < 1) return -1;
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
        ES_CUDA(cudaFuncSetAttribute(umma_phoneme_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PH_SMEM));
        ES_CUDA(cudaFuncSetAttribute(umma_phoneme_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PH_SMEM));
        attr_set = true;
    }
    PhonemeParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.N = N; p.n1 = n1; p.pool = pool; p.n_symbols = cfg.n_symbols;
    p.ids = ids; p.mask = mask; p.pitch_tgt = pitch_tgt; p.energy_tgt = energy_tgt; p.dur_tgt = dur_tgt;
    p.enc[0] = w.enc[0]; p.enc[1] = w.enc[1];
    p.fuse_u_h16 = w.fuse_u_h16; p.fuse_gb = w.fuse_gb; p.fuse_a0_h16 = w.fuse_a0_h16; p.fuse_c = w.fuse_c;
    p.pred[0] = w.pitch; p.pred[1] = w.energy; p.pred[2] = w.duration;
    p.pitch_pred = pitch_pred; p.energy_pred = energy_pred; p.dur_pred = dur_pred; p.fused4 = fused4;
    p.dur_int = dur_int; p.dur_cum = dur_cum; p.mel_len = mel_len;
    p.sc_xm1 = sc_xm1;
    p.scale_log2e = 1.4426950408889634f / sqrtf(32.f);
    p.err = err_flag;
    p.trace = g_ph_trace;
    const int grid = B < n_sm ? B : n_sm;
    ES_CUDA(launch_pdl(p.trace ? umma_phoneme_kernel<true> : umma_phoneme_kernel<false>, grid, NTHR, PH_SMEM, s, p));
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace es

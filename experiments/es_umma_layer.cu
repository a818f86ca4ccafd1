// Decoder layer for the 128-channel decoders (tiny), second generation: MelDecoder's
//     x = LN_l(tanh(Conv1x1(DWConv_k5(x))))  [ ; skip = LN_blk(x + skip) on the last layer of a block ]
// (layers/networks.py:281-283, :297-299) as ONE kernel per layer: tcgen05 + TMEM + a wide warp-specialised CTA.
//
// What the first generation (es_umma_dec.cu, 13 warps at 128 registers) taught -- measured, DESIGN.md section 5:
// the layer is bound by instruction ISSUE, not by HBM or the tensor pipe (issue slots 33-39 % busy, 3 warps per
// scheduler with MUFU / shared-memory / TMEM dependency chains), and with the epilogue removed the four producer
// warps alone still need 3.8 k cycles per 64-frame tile.  So this kernel is built for thread-level parallelism:
//
//   * 25 working warps: 16 epilogue (64 registers), 8 producers (104), 1 issuer -- setmaxnreg moves registers from the
//     epilogue and issue warpgroups to the producers, whose depthwise taps then live in registers for the whole kernel;
//   * 128-frame tiles (one M=128 tcgen05.mma per K step: half the tensor time per frame of M=64, half the
//     hand-offs), two 128-column TMEM accumulators, ONE A buffer (the producers' loads and depthwise conv of tile
//     i+1 overlap the GEMM of tile i; only their operand stores wait for it);
//   * no x ring and no TMA for activations: a producer lane owns 4 channels and slides a 12-row window of 16-byte
//     global loads (a warp-wide load is one full 512-byte row, 12 in flight per thread); the rows of the NEXT tile are
//     pulled into L2 (prefetch.global.L2) while the current one is processed, so the loads see L2 latency, not HBM's.
//     The depthwise conv runs straight out of registers, and the gathered first block (rows of the per-phoneme
//     projection table addressed through the frame -> row map, es_gather.cu) is the same code with an index load;
//   * the epilogue splits every accumulator row between TWO warps (columns 0-63 / 64-127; tcgen05.ld 16x256b.x8,
//     the mma-fragment layout: 4 threads per row segment): 32 values per thread instead of 64, twice the warps in
//     flight.  The LayerNorm statistics of the two halves meet in shared memory (one 64-thread named barrier per
//     statistic).  bias -> tanh -> LayerNorm [-> + skip -> LayerNorm] in registers;
//   * the OUTPUT leaves through shared memory and bulk async stores (TMA), not st.global: ncu showed the L1 data pipe
//     (LSU wavefronts, shared with the tensor core's operand fetches) to be the busiest unit of the layer, and a
//     global store costs it one wavefront per 32 bytes whatever the instruction width, a shared-memory store one per
//     128.  Each warp pair stages its 16 x 128 piece in its own padded buffer and one lane issues 16 row copies.
//
// Split-fp16 arithmetic as everywhere (es_umma.cuh): A and W as fp16 hi + lo, three MMAs per K step, fp32
// accumulation.  Shared memory: W (hi, lo) 64 KB resident, the A buffer 66 KB, 8 staging buffers of 8.3 KB, parameters.
// mbarriers: bar_w, bar_aready (8 producer warps), bar_mma[2] (accumulator full == the A buffer free),
// bar_tfree[2] (accumulator drained by its 8 epilogue warps).  Every wait is bounded (device error flag).
#include "es_common.cuh"
#include "es_kernels.cuh"
#include "es_umma.cuh"

namespace es {
namespace {

using namespace umma;

constexpr int TM = 128;                   // frames per tile (UMMA M)
constexpr int CK = 128;                   // channels (K and N)
constexpr int DWK = 5, HALO = 2;
constexpr int N_EPI = 16, N_PROD = 8;
constexpr int NTHR = (N_EPI + N_PROD + 4) * 32;          // 896: 4 + 2 warpgroups + one for the issue warp (3 of its warps idle)
constexpr int REG_EPI = 64, REG_PROD = 104, REG_MISC = 40;   // setmaxnreg (launch: 72 each): epilogue and issue warpgroups hand registers to the producers
constexpr uint32_t A_LBO = TM * 16 + 16;  // K-panel stride: [128 rows x 16 B] + 16 B pad (bank spread of the lane stores)
constexpr uint32_t A_PLANE = 16 * A_LBO;  // 33024: one fp16 plane (hi or lo), 16 K panels
constexpr uint32_t A_BUF = 2 * A_PLANE;   // 66048
constexpr uint32_t W_PLANE = CK * CK * 2; // 32768

constexpr uint32_t ST_LD = CK * 4 + 32;   // staging row stride (bytes): 512 + 32 -- the 4 rows x 32 B a half-warp touches with one
                                          // 8-byte fragment access then tile the 32 banks exactly
constexpr uint32_t ST_BUF = 16 * ST_LD;   // 8704: one 16-row piece

constexpr uint32_t OFF_A = 0;
constexpr uint32_t OFF_W = OFF_A + A_BUF;                     // 66048
constexpr uint32_t OFF_PAR = OFF_W + 2 * W_PLANE;             // bias (pre-scaled), ln g/b, ln2 g/b
constexpr uint32_t OFF_DW = OFF_PAR + 5 * CK * 4;             // depthwise taps + bias
constexpr uint32_t OFF_XCH = OFF_DW + 6 * CK * 4;             // LayerNorm partial sums: [2 LN][2 slots][4 q][2 halves][2 chunks][8 rows] float4
constexpr uint32_t OFF_ST = OFF_XCH + 2 * 2 * 4 * 2 * 2 * 8 * 16;   // output staging: one buffer per epilogue warp pair
constexpr uint32_t OFF_BAR = OFF_ST + 8 * ST_BUF;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256;
static_assert(OFF_W % 128 == 0, "operand alignment");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct LayerParams {
    int B, T;
    const float* X;              // [B,T,128], or the projection table when `src` gathers the input
    const float* dw_w;           // [5][128]
    const float* dw_b;           // [128]
    const void* w_h16;           // canonical split-fp16 weights [2][16][128][8]
    const float* bias;
    const float* ln_g; const float* ln_b;
    const float* res2;           // skip [B,T,128] (or the table when gather_res2), null on plain layers
    const float* ln2_g; const float* ln2_b;
    const int* src;              // frame -> table row map [B*T] (gathered first block) or null
    float* Y;                    // [B,T,128]
    int* err;
    long long* trace;            // -DES_LAYER_TRACE builds: clock64 stamps of CTA 0, [6 roles][16 tiles][8 events]
};

#ifdef ES_LAYER_TRACE
#define LTRACE(role, iter, ev)                                                                                 \
    do {                                                                                                       \
        if (p.trace && blockIdx.x == 0 && (iter) < 16) p.trace[((role) * 16 + (iter)) * 8 + (ev)] = clock64(); \
    } while (0)
#else
#define LTRACE(role, iter, ev) do { } while (0)
#endif

constexpr float kTanhScaleL = 2.8853900817779268f;           // 2 log2(e): tanh(x) = 1 - 2 / (1 + 2^(x * 2 log2 e))
__device__ __forceinline__ float ex2a(float x) { float e; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x)); return e; }
__device__ __forceinline__ float rcpa(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ uint32_t a_off(int row, int lane) {
    return (uint32_t)(lane >> 1) * A_LBO + (uint32_t)row * 16u + (uint32_t)(lane & 1) * 8u;
}

// quad reduction of (sum, sum of squares) packed as (row0, row1) pairs -> rstd and -mean*rstd per row
__device__ __forceinline__ void row_stats(f32x2 s0, f32x2 q0, f32x2 s1, f32x2 q1, float& r0, float& n0, float& r1, float& n1) {
    const float2 a0 = up2(s0), b0 = up2(q0), a1 = up2(s1), b1 = up2(q1);
    quad_stats(a0.x + a0.y, b0.x + b0.y, 1.f / CK, r0, n0);
    quad_stats(a1.x + a1.y, b1.x + b1.y, 1.f / CK, r1, n1);
}

// One 16-row x 64-column piece of an accumulator (this warp's TMEM lanes, its column half), held in registers.
// Fragment layout of tcgen05.ld 16x256b.x8: r[4j + 2i + b] = row (lane/4 + 8i), column 8j + 2(lane%4) + b.
// xmine / xpeer: this warp's and the partner warp's float4 slots (sum0, sumsq0, sum1, sumsq1) of the piece's 8 row
// pairs; bar_id: the pair's named barrier.
__device__ __forceinline__ void pair_stats(f32x2 s0, f32x2 q0, f32x2 s1, f32x2 q1, float4* xmine, const float4* xpeer,
                                           int bar_id, int t4, int tr, float& r0, float& n0, float& r1, float& n1) {
    const float2 a0 = up2(s0), b0 = up2(q0), a1 = up2(s1), b1 = up2(q1);
    float S0 = a0.x + a0.y, Q0 = b0.x + b0.y, S1 = a1.x + a1.y, Q1 = b1.x + b1.y;
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
        S0 += __shfl_xor_sync(0xffffffffu, S0, o); Q0 += __shfl_xor_sync(0xffffffffu, Q0, o);
        S1 += __shfl_xor_sync(0xffffffffu, S1, o); Q1 += __shfl_xor_sync(0xffffffffu, Q1, o);
    }
    if (t4 == 0) xmine[tr] = make_float4(S0, Q0, S1, Q1);
    named_bar_sync(bar_id, 64);
    const float4 o = xpeer[tr];
    constexpr float inv_n = 1.f / CK;
    const float m0 = (S0 + o.x) * inv_n, m1 = (S1 + o.z) * inv_n;
    r0 = rsqrtf(fmaf(Q0 + o.y, inv_n, -m0 * m0) + kLnEpsU);
    r1 = rsqrtf(fmaf(Q1 + o.w, inv_n, -m1 * m1) + kLnEpsU);
    n0 = -m0 * r0;
    n1 = -m1 * r1;
}

#ifdef ES_EXP_L_NOSTORE       // timing experiment: the stores become (practically never taken) data-dependent stores
#define ES_STORE_OK(ok, v) ((ok) && (v) == 0x7fc123457fc12345ull)
#else
#define ES_STORE_OK(ok, v) (ok)
#endif
// par: parameters of this warp's column half (permuted layout, see the kernel prologue); skip0/skip1: row pointers
// already offset to the column half
// stage: this warp's column half of the pair's staging buffer (col_off bytes into its rows); issuer: the one lane of the
// pair that drives the TMA; yrow: global address of the piece's first row; n_rows: valid rows of the piece
// Block-end layers: the piece's 16 skip rows are bulk-loaded (TMA) into the SAME staging buffer while the tanh /
// LayerNorm math runs (skip_rows: their global address when the rows are contiguous, or srcmap != null: one table row
// per frame through the frame -> row map); every thread then reads and later overwrites exactly its own elements.
template <bool RES2, bool GS>
__device__ __forceinline__ bool epilogue_piece(uint32_t tpiece, const float* par, const float* skip_rows, const int* srcmap,
                                               uint32_t skip_bar, uint32_t skip_parity, int t4, int tr, float4* xmine,
                                               const float4* xpeer, int bar_id, uint8_t* stage, uint32_t col_off, bool issuer,
                                               float* yrow, int n_rows) {
    uint32_t r[32];
    bool ok = true;
    if (issuer) {
        bulk_wait_read0();                    // the previous piece has left the staging buffer
        if (RES2 && n_rows > 0) {
            const uint32_t dst = smem_u32(stage) - col_off;
            mbar_arrive_expect_tx(skip_bar, (uint32_t)n_rows * CK * 4u);
            for (int rr = 0; rr < n_rows; ++rr) {
                const float* srow = GS ? skip_rows + (size_t)__ldg(srcmap + rr) * CK : skip_rows + (size_t)rr * CK;
                bulk_g2s(dst + (uint32_t)rr * ST_LD, srow, CK * 4, skip_bar);
            }
        }
    }
    tmem_ld_16x256b_x8(tpiece, r);
    tmem_ld_wait();
    const ulonglong2* par2 = reinterpret_cast<const ulonglong2*>(par) + t4;      // [jp * 4] -> (j = 2jp pair, j = 2jp+1 pair)
    f32x2 v[16];                          // v[2j + i]: row i, columns 8j + 2 t4, +1
    f32x2 s0 = 0ull, q0 = 0ull, s1 = 0ull, q1 = 0ull;
    {
        const f32x2 cs = pk2(kTanhScaleL, kTanhScaleL), one2 = pk2(1.f, 1.f), mtwo2 = pk2(-2.f, -2.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const ulonglong2 b4 = par2[(j >> 1) * 4];
            const f32x2 bb = (j & 1) ? b4.y : b4.x;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float2 a = up2(fma2(pk2u(r[4 * j + 2 * i], r[4 * j + 2 * i + 1]), cs, bb));
                const float2 d = up2(add2(pk2(ex2a(a.x), ex2a(a.y)), one2));
                const f32x2 t = fma2(mtwo2, pk2(rcpa(d.x), rcpa(d.y)), one2);
                v[2 * j + i] = t;
                if (i == 0) { s0 = add2(s0, t); q0 = fma2(t, t, q0); } else { s1 = add2(s1, t); q1 = fma2(t, t, q1); }
            }
        }
    }
    float r0, n0, r1, n1;
    pair_stats(s0, q0, s1, q1, xmine, xpeer, bar_id, t4, tr, r0, n0, r1, n1);
    {
        const f32x2 R0 = pk2(r0, r0), N0 = pk2(n0, n0), R1 = pk2(r1, r1), N1 = pk2(n1, n1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const ulonglong2 g4 = par2[32 + (j >> 1) * 4], e4 = par2[64 + (j >> 1) * 4];
            const f32x2 gg = (j & 1) ? g4.y : g4.x, be = (j & 1) ? e4.y : e4.x;
            v[2 * j] = fma2(fma2(v[2 * j], R0, N0), gg, be);
            v[2 * j + 1] = fma2(fma2(v[2 * j + 1], R1, N1), gg, be);
        }
    }
    if (RES2) {
        // + skip (from the staging buffer), second statistics (the other exchange slots), block LayerNorm
        s0 = 0ull; q0 = 0ull; s1 = 0ull; q1 = 0ull;
        if (n_rows > 0 && !mbar_wait(skip_bar, skip_parity)) ok = false;
        const bool ok0 = tr < n_rows, ok1 = tr + 8 < n_rows;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t c = (uint32_t)(8 * j + 2 * t4) * 4u;
            const f32x2 k0 = ok0 ? *reinterpret_cast<const f32x2*>(stage + (uint32_t)tr * ST_LD + c) : 0ull;
            const f32x2 k1 = ok1 ? *reinterpret_cast<const f32x2*>(stage + (uint32_t)(tr + 8) * ST_LD + c) : 0ull;
            v[2 * j] = add2(v[2 * j], k0);
            v[2 * j + 1] = add2(v[2 * j + 1], k1);
            s0 = add2(s0, v[2 * j]); q0 = fma2(v[2 * j], v[2 * j], q0);
            s1 = add2(s1, v[2 * j + 1]); q1 = fma2(v[2 * j + 1], v[2 * j + 1], q1);
        }
        constexpr int LN_STRIDE = 2 * 4 * 2 * 2 * 8;          // float4 slots of one statistic
        pair_stats(s0, q0, s1, q1, xmine + LN_STRIDE, xpeer + LN_STRIDE, bar_id, t4, tr, r0, n0, r1, n1);
        const f32x2 R0 = pk2(r0, r0), N0 = pk2(n0, n0), R1 = pk2(r1, r1), N1 = pk2(n1, n1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const ulonglong2 g4 = par2[96 + (j >> 1) * 4], e4 = par2[128 + (j >> 1) * 4];
            const f32x2 gg = (j & 1) ? g4.y : g4.x, be = (j & 1) ? e4.y : e4.x;
            v[2 * j] = fma2(fma2(v[2 * j], R0, N0), gg, be);
            v[2 * j + 1] = fma2(fma2(v[2 * j + 1], R1, N1), gg, be);
        }
    }
    // Output: fragment -> this pair's staging buffer (8-byte stores, 8 rows x 32 B at a 544-byte row stride: two
    // wavefronts, the minimum for 256 bytes), then one lane hands the 16 rows to the TMA.  The previous piece's copies
    // have finished READING the buffer: the issuing lane waited for that before the statistics barrier above.
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t c = (uint32_t)(8 * j + 2 * t4) * 4u;
        *reinterpret_cast<f32x2*>(stage + (uint32_t)tr * ST_LD + c) = v[2 * j];
        *reinterpret_cast<f32x2*>(stage + (uint32_t)(tr + 8) * ST_LD + c) = v[2 * j + 1];
    }
    fence_proxy_async_smem();
    named_bar_sync(bar_id, 64);
#ifndef ES_EXP_L_NOSTORE
    if (issuer) {
        const uint32_t src = smem_u32(stage) - col_off;                 // (stage points at this warp's column half)
        for (int rr = 0; rr < n_rows; ++rr) bulk_s2g(yrow + (size_t)rr * CK, src + (uint32_t)rr * ST_LD, CK * 4);
        bulk_commit();
    }
#endif
    return ok;
}

// RES2: block-end layer.  GX: input rows come from the table through p.src.  GS: so do the skip rows.
template <bool RES2, bool GX, bool GS>
__global__ void __launch_bounds__(NTHR, 1)
umma_layer_kernel(const LayerParams p) {
    static_assert(!GS || RES2, "gathered skip rows need the block-end epilogue");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float* par = reinterpret_cast<float*>(smem + OFF_PAR);
    const uint32_t bar0 = smem_u32(smem + OFF_BAR);
    const uint32_t bar_w = bar0;                       //     weights landed
    const uint32_t bar_aready = bar0 + 8;              //     A buffer written by the 8 producer warps
    const uint32_t bar_mma = bar0 + 24;                // [2] accumulator full / its A buffer free
    const uint32_t bar_tfree = bar0 + 40;              // [2] accumulator drained
    const uint32_t bar_skip = bar0 + 56;               // [8] skip rows of a piece landed in the pair's staging buffer (block-end layers)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 128);

    const int tiles_per_utt = (p.T + TM - 1) / TM;
    const int n_tiles = p.B * tiles_per_utt;

    // ---- one-time setup ---------------------------------------------------------------------
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 256);          // two 128-column accumulators
    if (tid == 32) {
        mbar_init(bar_w, 1);
        mbar_init(bar_aready, N_PROD);
        for (int k = 0; k < 2; ++k) { mbar_init(bar_mma + 8 * k, 1); mbar_init(bar_tfree + 8 * k, 8); }
        for (int k = 0; k < 8; ++k) mbar_init(bar_skip + 8 * k, 1);
        fence_mbar_init();
        // the weights do not depend on the predecessor kernel: their bulk load starts before pdl_wait()
        mbar_arrive_expect_tx(bar_w, 2 * W_PLANE);
        bulk_g2s(smem_u32(smem + OFF_W), p.w_h16, W_PLANE, bar_w);
        bulk_g2s(smem_u32(smem + OFF_W) + W_PLANE, reinterpret_cast<const uint8_t*>(p.w_h16) + W_PLANE, W_PLANE, bar_w);
    }
    // epilogue parameters, permuted to the fragment layout: thread t4 of a quad works on columns 8j + 2 t4, +1; slot
    // [jp][t4] (16 bytes) holds them for j = 2 jp and 2 jp + 1, so one LDS.128 serves two column groups
    for (int i = tid; i < CK; i += NTHR) {
        const int j = i >> 3, t4q = (i >> 1) & 3, e = i & 1;
        const int dst = ((j >> 1) * 4 + t4q) * 4 + (j & 1) * 2 + e;
        par[dst] = __ldg(p.bias + i) * kTanhScaleL;                // pre-scaled for the tanh argument
        par[128 + dst] = __ldg(p.ln_g + i);
        par[256 + dst] = __ldg(p.ln_b + i);
        par[384 + dst] = RES2 ? __ldg(p.ln2_g + i) : 0.f;
        par[512 + dst] = RES2 ? __ldg(p.ln2_b + i) : 0.f;
    }
    {
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

    if (warp >= N_EPI + N_PROD) {
        // =========================================================================== issue warp (+ 3 idle warps)
        setmaxnreg_dec<REG_MISC>();
        if (warp == N_EPI + N_PROD) {
        const uint32_t idesc = make_idesc_f16(TM, CK);
        const uint32_t lbo_b = (uint32_t)CK * 16u;
        const uint64_t dbh0 = make_smem_desc(smem_u32(smem + OFF_W), lbo_b, 128u);
        const uint64_t dbl0 = make_smem_desc(smem_u32(smem + OFF_W) + W_PLANE, lbo_b, 128u);
        const uint64_t db_step = (uint64_t)((2u * lbo_b) >> 4);
        const uint64_t da_step = (uint64_t)((2u * A_LBO) >> 4);
        const bool elected = elect_one();
        if (!mbar_wait(bar_w, 0)) failed = true;
        __syncwarp();
        int i = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
            const int slot = i & 1;
            if (elected) LTRACE(0, i, 0);
            if (!mbar_wait(bar_aready, i & 1)) failed = true;                             // A operand of tile i written
            if (elected) LTRACE(0, i, 1);
            if (i >= 2 && !mbar_wait(bar_tfree + 8 * slot, ((i >> 1) - 1) & 1)) failed = true;   // accumulator drained
            if (elected) LTRACE(0, i, 2);
            tc_fence_after_sync();
            const uint32_t a_base = smem_u32(smem + OFF_A);
            const uint64_t dah0 = make_smem_desc(a_base, A_LBO, 128u);
            const uint64_t dal0 = make_smem_desc(a_base + A_PLANE, A_LBO, 128u);
            const uint32_t acc = tmem + (uint32_t)(slot * CK);
#pragma unroll
            for (int k = 0; k < CK / 16; ++k) {
                const uint64_t da = (uint64_t)k * da_step, db = (uint64_t)k * db_step;
                if (elected) {
                    mma_f16_ss(acc, dah0 + da, dbh0 + db, idesc, k > 0 ? 1u : 0u);
                    mma_f16_ss(acc, dah0 + da, dbl0 + db, idesc, 1u);
                    mma_f16_ss(acc, dal0 + da, dbh0 + db, idesc, 1u);
                }
            }
            if (elected) { mma_commit(bar_mma + 8 * slot); LTRACE(0, i, 3); }
            __syncwarp();
        }
        }
    } else if (warp >= N_EPI) {
        // =========================================================================== producers
        setmaxnreg_inc<REG_PROD>();
        const int pw = warp - N_EPI;                                   // tile rows 16pw .. 16pw+15
        // depthwise taps + bias of this lane's 4 channels: registers for the whole kernel (re-reading them from shared
        // memory per row was a quarter of the L1 data-pipe wavefronts of this kernel)
        ulonglong2 wdw[DWK], bdw;
        {
            const ulonglong2* dwp = reinterpret_cast<const ulonglong2*>(smem + OFF_DW);   // [5 taps + bias][32 lanes]
#pragma unroll
            for (int t = 0; t < DWK; ++t) wdw[t] = dwp[t * 32 + lane];
            bdw = dwp[DWK * 32 + lane];
        }
        int i = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++i) {
            const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TM;
            uint8_t* a_hi = smem + OFF_A;
            const float* xb = p.X + (size_t)b * p.T * CK + 4 * lane;
            const int* sp = GX ? p.src + (size_t)b * p.T : nullptr;
            const bool tr_on = lane == 0 && (pw == 0 || pw == 7);
            const int trole = pw == 0 ? 1 : 2;
            if (tr_on) LTRACE(trole, i, 0);
            if (!GX) {
                // pull the 20 rows this warp reads from the NEXT tile (80 lines of 128 B) into L2 now: its loads will
                // find them there instead of paying HBM latency with a register window as the only buffer
                const int nt = tile + (int)gridDim.x;
                if (nt < n_tiles) {
                    const int nb = nt / tiles_per_utt, nt0 = (nt - nb * tiles_per_utt) * TM;
                    const float* nx = p.X + (size_t)nb * p.T * CK;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int line = lane + 32 * k;
                        const int t = nt0 + 16 * pw - HALO + (line >> 2);
                        if (line < 80 && t >= 0 && t < p.T) prefetch_l2(nx + (size_t)t * CK + (line & 3) * 32);
                    }
                }
            }
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int r0 = 16 * pw + 8 * pass;
                // 12-row window: frames t0 + r0 - 2 .. t0 + r0 + 9, zero outside the utterance (Conv1d zero padding)
                ulonglong2 win[8 + 2 * HALO];
#pragma unroll
                for (int k = 0; k < 8 + 2 * HALO; ++k) {
                    const int t = t0 + r0 - HALO + k;
#ifdef ES_EXP_L_NOLOAD     // timing experiment (never defined in the product build): no activation loads
                    const bool in = false; (void)t;
#else
                    const bool in = t >= 0 && t < p.T;
#endif
                    if (GX) {
                        const int row = in ? __ldg(sp + t) : 0;
                        win[k] = in ? __ldg(reinterpret_cast<const ulonglong2*>(p.X + (size_t)row * CK + 4 * lane)) : make_ulonglong2(0ull, 0ull);
                    } else {
                        win[k] = in ? __ldg(reinterpret_cast<const ulonglong2*>(xb + (size_t)t * CK)) : make_ulonglong2(0ull, 0ull);
                    }
                }
                if (tr_on) LTRACE(trole, i, 1 + 3 * pass);          // loads issued
                // A buffer free?  (the GEMM of tile i-1 has read it; this tile's loads are already in flight)
                if (pass == 0 && i >= 1 && !mbar_wait(bar_mma + 8 * ((i - 1) & 1), ((i - 1) >> 1) & 1)) failed = true;
                if (tr_on) LTRACE(trole, i, 2 + 3 * pass);          // A free
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    ulonglong2 o = bdw;
#pragma unroll
                    for (int t = 0; t < DWK; ++t) {
                        o.x = fma2(wdw[t].x, win[r + t].x, o.x);
                        o.y = fma2(wdw[t].y, win[r + t].y, o.y);
                    }
                    uint2 hi, lo;
                    split4(o, hi, lo);
                    const uint32_t off = a_off(r0 + r, lane);
                    *reinterpret_cast<uint2*>(a_hi + off) = hi;
                    *reinterpret_cast<uint2*>(a_hi + A_PLANE + off) = lo;
                }
                if (tr_on) LTRACE(trole, i, 3 + 3 * pass);          // pass stored
            }
            fence_proxy_async_smem();                  // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_aready);
            if (tr_on) LTRACE(trole, i, 7);
        }
    } else {
        // =========================================================================== epilogue
        // warp = 8 slot + 4 half + q: accumulator served, column half (0-63 / 64-127), TMEM lane quarter (rows 32q .. 32q+31)
        setmaxnreg_dec<REG_EPI>();
        const int slot = warp >> 3, half = (warp >> 2) & 1, q = warp & 3;
        const int t4 = lane & 3, tr = lane >> 2;
        const int bar_id = 1 + 4 * slot + q;                  // named barrier of the two warps sharing these rows
        float4* xch = reinterpret_cast<float4*>(smem + OFF_XCH);
        const float* parh = par + 64 * half;
        int n_skip = 0;                                       // skip loads this pair has waited for (mbarrier phase)
        int i = slot;
        for (int tile = blockIdx.x + slot * gridDim.x; tile < n_tiles; tile += 2 * gridDim.x, i += 2) {
            const int b = tile / tiles_per_utt, t0 = (tile - b * tiles_per_utt) * TM;
            const int rows_valid = min(TM, p.T - t0);
            const bool tr_on = lane == 0 && (warp == 0 || warp == 4 || warp == 8);
            const int trole = warp == 0 ? 3 : warp == 4 ? 4 : 5;
            if (tr_on) LTRACE(trole, i, 0);
            if (!mbar_wait(bar_mma + 8 * slot, (i >> 1) & 1)) failed = true;
            if (tr_on) LTRACE(trole, i, 1);
            tc_fence_after_sync();
#ifdef ES_EXP_L_NOEPI        // timing experiment: the epilogue only hands the accumulator back
            if (false)
#endif
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                const size_t gp = (size_t)b * p.T + t0 + 32 * q + 16 * c;            // first frame of the piece
                const int n_rows = max(0, min(16, rows_valid - (32 * q + 16 * c)));
                const uint32_t tpiece = tmem + ((uint32_t)(32 * q + 16 * c) << 16) + (uint32_t)(slot * CK + 64 * half);
                float4* xmine = xch + ((((slot * 4 + q) * 2 + half) * 2 + c) * 8);
                const float4* xpeer = xch + ((((slot * 4 + q) * 2 + (half ^ 1)) * 2 + c) * 8);
                const int pair = 4 * slot + q;
                if (!epilogue_piece<RES2, GS>(tpiece, parh, RES2 ? (GS ? p.res2 : p.res2 + gp * CK) : nullptr, GS ? p.src + gp : nullptr,
                                              bar_skip + 8 * pair, (uint32_t)(n_skip & 1), t4, tr, xmine, xpeer, bar_id,
                                              smem + OFF_ST + (uint32_t)pair * ST_BUF + 256u * half, 256u * half,
                                              half == 0 && lane == 0, p.Y + gp * CK, n_rows)) failed = true;
                if (RES2 && n_rows > 0) ++n_skip;
                if (tr_on) LTRACE(trole, i, 2 + c);
            }
            // every tcgen05.ld of this warp has completed (tmem_ld_wait inside): the accumulator may be overwritten
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tfree + 8 * slot);
            if (tr_on) LTRACE(trole, i, 4);
        }
        if (half == 0 && lane == 0) bulk_wait0();             // this lane's bulk stores have been written
    }

    if (failed) atomicExch(p.err, 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

template <bool RES2, bool GX, bool GS>
int launch_layer(const LayerParams& p, int grid, cudaStream_t s) {
    static PerDeviceSlot<bool> attr_once;
    bool& attr_set = attr_once.get();                   // function attributes are per device
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(umma_layer_kernel<RES2, GX, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    ES_CUDA(launch_pdl(umma_layer_kernel<RES2, GX, GS>, grid, NTHR, SMEM_BYTES, s, p));
    ES_LAUNCH_OK();
    return 0;
}

long long* g_layer_trace = nullptr;
int g_layer_trace_pick = 0, g_layer_trace_count = 0;

}  // namespace

// debug (-DES_LAYER_TRACE builds): the `pick`-th layer launch after this call stamps clock64() per role / tile / event
void umma_layer_set_trace(long long* buf, int pick) {
    g_layer_trace = buf;
    g_layer_trace_pick = pick;
    g_layer_trace_count = 0;
}

bool umma_layer_supported(int C, int dw_k) { return C == CK && dw_k == DWK; }

// One decoder layer (networks.py:281-283, :297-299).  res2 != null: block-end layer (skip add + block LayerNorm).
// src != null: frame -> table row map; gather_x: X is the table, gather_res2: res2 is the table.
int launch_umma_layer(int B, int T, const float* X, const float* dw_w, const float* dw_b, const void* w_h16,
                      const float* bias, const float* ln_g, const float* ln_b, const float* res2, const float* ln2_g,
                      const float* ln2_b, const int* src, int gather_x, int gather_res2, float* Y, cudaStream_t s) {
    ES_CHECK(X && dw_w && dw_b && w_h16 && bias && ln_g && ln_b && Y, "null tensor");
    ES_CHECK(!res2 || (ln2_g && ln2_b), "the skip path needs the block LayerNorm");
    ES_CHECK(!(gather_x || gather_res2) || src, "gathered rows need the frame -> row map");
    ES_CHECK(!gather_res2 || res2, "gathered skip without a table");
    ES_CHECK(B >= 1 && T >= 1, "empty batch");
    int* err_flag = umma_err_flag();
    ES_CHECK(err_flag, "cannot allocate the device error flag");
    static PerDeviceSlot<int> n_sm_once;
    int& n_sm = n_sm_once.get();
    if (!n_sm) {
        int dev = 0;
        ES_CUDA(cudaGetDevice(&dev));
        ES_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    LayerParams p;
    p.B = B; p.T = T; p.X = X; p.dw_w = dw_w; p.dw_b = dw_b; p.w_h16 = w_h16; p.bias = bias;
    p.ln_g = ln_g; p.ln_b = ln_b; p.res2 = res2; p.ln2_g = ln2_g; p.ln2_b = ln2_b; p.src = src; p.Y = Y; p.err = err_flag;
    p.trace = (g_layer_trace && g_layer_trace_count++ == g_layer_trace_pick) ? g_layer_trace : nullptr;
    const int n_tiles = B * ((T + TM - 1) / TM);
    const int grid = n_tiles < n_sm ? n_tiles : n_sm;
    if (!res2) return gather_x ? launch_layer<false, true, false>(p, grid, s) : launch_layer<false, false, false>(p, grid, s);
    if (gather_x) return gather_res2 ? launch_layer<true, true, true>(p, grid, s) : launch_layer<true, true, false>(p, grid, s);
    return gather_res2 ? launch_layer<true, false, true>(p, grid, s) : launch_layer<true, false, false>(p, grid, s);
}

}  // namespace es

// Tensor-core GEMM of the training step: C[M,N] = op(A)[M,K] op(B)[K,N] (+ bias) on tcgen05 with split 16-bit operands.
//
// The training path (es_train_ops.cu) spends two thirds of its time in three GEMM shapes per decoder layer, all with
// ~1e5 rows: Y = X W^T (forward), dX = dY W (input gradient) and dW = dY^T X (weight gradient, contraction over the
// rows).  fp32 has no tensor-core path, so each fp32 operand value v is split into two 16-bit halves hi = rn16(v),
// lo = rn16(v - hi) and a product is three MMAs (hi.hi + hi.lo + lo.hi) accumulated in fp32 in TMEM -- the scheme of
// the inference kernels (es_umma_dec.cu).  What differs here:
//   * operands arrive as fp32 in any of the four transpose combinations; the producer threads read them with the
//     access pattern that coalesces for that orientation and write the canonical K-major core-matrix layout, so one
//     kernel serves forward, dX and dW without a transposed copy in HBM;
//   * products with a GRADIENT operand use a bf16 split instead of fp16: gradients of a mean loss over ~1e7 elements sit
//     around 1e-7, below fp16's normal range, where an fp16 split would lose them; bf16 keeps fp32's exponent (2 x 8
//     mantissa bits: relative 2^-17 per element, far inside what training needs and inside the parity tests' 2e-4).
//     kind::f16 does not accept one fp16 and one bf16 operand (the instruction faults), so both operands of such a
//     product are split as bf16;
//   * the contraction over rows (dW) is split across CTAs (split-K), each writing a partial tile; the caller adds the
//     partials in a fixed order (es_t_colsum), so weight gradients are deterministic.
//
// One CTA = one 128-row x N<=256 tile, K streamed in chunks of 32 through a 2-stage ring: all 128 threads convert and
// stage chunk c while the tensor core works on chunk c-1; a stage is recycled when the commit of the MMAs that read it
// arrives on its mbarrier.  2-3 CTAs are resident per SM, which is what hides the global-load latency of the producers.
#include "es_common.cuh"
#include "es_kernels.cuh"
#include "es_umma.cuh"

#include <cuda_bf16.h>

namespace es {
namespace {

using namespace umma;

constexpr int GM = 128;      // tile rows (TMEM lanes)
constexpr int KC = 32;       // K per stage: 4 core-matrix columns, 2 MMA k-steps
constexpr int STAGES = 2;

struct TcGemmParams {
    const float* A; const float* B; float* C; const float* bias;
    int M, N, K, lda, ldb, ldc, ta, tb;
    int k_chunk;             // > 0: blockIdx.z owns k in [z k_chunk, ...) and writes C + z * sc
    long long sa, sb, sc;    // k_chunk == 0: blockIdx.z is a batch index, operands advance by these strides
    int wide;                // an operand holds gradients: both operands are split as bf16 pairs instead of fp16 pairs
    int ntp;                 // N tile padded to 16
    int* err;
};

__device__ __forceinline__ void split8_wide(const float (&v)[8], uint4& hi, uint4& lo) {
    __nv_bfloat162 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        const float2 back = __bfloat1622float2(hh);
        h[i] = hh;
        l[i] = __floats2bfloat162_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);
    }
    hi = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                    *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
    lo = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]),
                    *reinterpret_cast<uint32_t*>(&l[2]), *reinterpret_cast<uint32_t*>(&l[3]));
}

// Stage `rows_pad` rows x KC of one operand.  Element (r, k) of the operand is src[r * ld + k] (rowmajor) or
// src[k * ld + r] (transposed storage); r in [r0, r0 + rows_pad) valid below r_end, k in [k0, k0 + KC) valid below k_end.
// The loads of UNROLL items (2 x 16 bytes or 8 x 4 bytes each) are all issued before the first conversion: with 12
// resident warps per SM the kernel lives on memory-level parallelism per thread, not on occupancy (the first version
// converted item by item and sat at 11 % of DRAM bandwidth, stalled on long_scoreboard; profiles/r02_ncu_train_gemm.txt).
template <bool WIDE>
__device__ __forceinline__ void stage_operand(const float* __restrict__ src, int ld, bool transposed, int r0, int r_end, int k0, int k_end,
                                              int rows_pad, uint8_t* hi_base, uint8_t* lo_base, int tid, bool vec_ok) {
    constexpr int UNROLL = 4;
    const int items = rows_pad * (KC / 8);
    for (int base = 0; base < items; base += UNROLL * 128) {
        float v[UNROLL][8];
        uint32_t off[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int idx = base + u * 128 + tid;
            int r, kc;
            if (transposed) { r = idx % rows_pad; kc = idx / rows_pad; }      // lanes along r: coalesced for [k][r] storage
            else            { kc = idx & 3; r = idx >> 2; }                   // 4 lanes cover 128 contiguous bytes of a row
            off[u] = idx < items ? canon_off(r, kc, rows_pad) : 0xffffffffu;
            const int gr = r0 + r, gk = k0 + kc * 8;
            if (idx < items && gr < r_end && !transposed && vec_ok && gk + 8 <= k_end) {
                const float4* p = reinterpret_cast<const float4*>(src + (size_t)gr * ld + gk);
                const float4 a = __ldg(p), b = __ldg(p + 1);
                v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = b.x; v[u][5] = b.y; v[u][6] = b.z; v[u][7] = b.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int k = gk + i;
                    v[u][i] = (idx < items && gr < r_end && k < k_end)
                                  ? __ldg(transposed ? src + (size_t)k * ld + gr : src + (size_t)gr * ld + k) : 0.f;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (off[u] == 0xffffffffu) continue;
            uint4 hi, lo;
            if (WIDE) split8_wide(v[u], hi, lo); else split8(v[u], hi, lo);
            *reinterpret_cast<uint4*>(hi_base + off[u]) = hi;
            *reinterpret_cast<uint4*>(lo_base + off[u]) = lo;
        }
    }
}

__global__ void __launch_bounds__(128)
t_gemm_tc_kernel(const TcGemmParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t free_bar[STAGES];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * GM, n0 = blockIdx.y * p.ntp;
    const int n_end = min(p.N, n0 + p.ntp);
    int k_lo = 0, k_hi = p.K;
    const float* Ap = p.A;
    const float* Bp = p.B;
    float* C = p.C + (long long)blockIdx.z * p.sc;
    if (p.k_chunk > 0) {
        k_lo = blockIdx.z * p.k_chunk;
        k_hi = min(p.K, k_lo + p.k_chunk);
    } else {
        Ap += (long long)blockIdx.z * p.sa;
        Bp += (long long)blockIdx.z * p.sb;
    }
    const uint32_t a_bytes = GM * KC * 2, b_bytes = (uint32_t)p.ntp * KC * 2, stage_bytes = 2 * a_bytes + 2 * b_bytes;
    const uint32_t ncols = p.ntp <= 32 ? 32u : p.ntp <= 64 ? 64u : p.ntp <= 128 ? 128u : 256u;

    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), ncols);
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&free_bar[s]), 1);
        fence_mbar_init();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    const bool a_vec = !p.ta && (p.lda & 3) == 0 && ((reinterpret_cast<uintptr_t>(Ap) & 15) == 0) && (k_lo & 3) == 0;
    const bool b_vec = p.tb && (p.ldb & 3) == 0 && ((reinterpret_cast<uintptr_t>(Bp) & 15) == 0) && (k_lo & 3) == 0;
    // a/b_format: 0 = F16, 1 = BF16 (bits [7,10) and [10,13) of the kind::f16 instruction descriptor)
    const uint32_t idesc = make_idesc_f16(GM, p.ntp) | (p.wide ? ((1u << 7) | (1u << 10)) : 0u);
    const int nkc = (k_hi - k_lo + KC - 1) / KC;
    bool failed = false;

    for (int c = 0; c < nkc; ++c) {
        const int s = c & 1;
        if (c >= STAGES && !mbar_wait(smem_u32(&free_bar[s]), (uint32_t)(((c >> 1) - 1) & 1))) failed = true;
        uint8_t* st = smem + (size_t)s * stage_bytes;
        uint8_t* a_hi = st; uint8_t* a_lo = st + a_bytes; uint8_t* b_hi = st + 2 * a_bytes; uint8_t* b_lo = b_hi + b_bytes;
        const int k0 = k_lo + c * KC;
        // op(A)(m, k): stored [m][k] (ta = 0) or [k][m] (ta = 1)
        if (p.wide) stage_operand<true>(Ap, p.lda, p.ta != 0, m0, p.M, k0, k_hi, GM, a_hi, a_lo, tid, a_vec);
        else          stage_operand<false>(Ap, p.lda, p.ta != 0, m0, p.M, k0, k_hi, GM, a_hi, a_lo, tid, a_vec);
        // op(B)(k, n): the canonical tile holds rows n, columns k; stored [n][k] (tb = 1) or [k][n] (tb = 0)
        if (p.wide) stage_operand<true>(Bp, p.ldb, p.tb == 0, n0, n_end, k0, k_hi, p.ntp, b_hi, b_lo, tid, b_vec);
        else          stage_operand<false>(Bp, p.ldb, p.tb == 0, n0, n_end, k0, k_hi, p.ntp, b_hi, b_lo, tid, b_vec);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
            const uint32_t lbo_a = GM * 16u, lbo_b = (uint32_t)p.ntp * 16u, sbo = 128u;
#pragma unroll
            for (int ks = 0; ks < KC / 16; ++ks) {
                const uint32_t ao = (uint32_t)(2 * ks) * lbo_a, bo = (uint32_t)(2 * ks) * lbo_b;
                const uint64_t dah = make_smem_desc(smem_u32(a_hi) + ao, lbo_a, sbo);
                const uint64_t dal = make_smem_desc(smem_u32(a_lo) + ao, lbo_a, sbo);
                const uint64_t dbh = make_smem_desc(smem_u32(b_hi) + bo, lbo_b, sbo);
                const uint64_t dbl = make_smem_desc(smem_u32(b_lo) + bo, lbo_b, sbo);
                mma_f16_ss(tmem, dah, dbh, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                mma_f16_ss(tmem, dah, dbl, idesc, 1u);
                mma_f16_ss(tmem, dal, dbh, idesc, 1u);
            }
            mma_commit(smem_u32(&free_bar[s]));
        }
    }
    // the last commit covers every MMA issued before it
    if (nkc > 0 && !mbar_wait(smem_u32(&free_bar[(nkc - 1) & 1]), (uint32_t)(((nkc - 1) >> 1) & 1))) failed = true;
    tc_fence_after_sync();

    const int row = m0 + warp * 32 + lane;
    const bool vec_store = (p.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    for (int c0 = 0; c0 < p.ntp; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        if (row < p.M && nkc > 0) {
            float* dst = C + (size_t)row * p.ldc + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const int col = n0 + c0 + j;
                if (col >= n_end) break;
                float v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[j + i]) + ((p.bias && col + i < n_end) ? __ldg(p.bias + col + i) : 0.f);
                if (vec_store && col + 4 <= n_end) *reinterpret_cast<float4*>(dst + j) = make_float4(v[0], v[1], v[2], v[3]);
                else
                    for (int i = 0; i < 4 && col + i < n_end; ++i) dst[j + i] = v[i];
            }
        }
    }
    if (failed && lane == 0) atomicExch(p.err, 1);
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

PerDeviceSlot<int> g_tc_enabled_init;      // 0: not read yet, 1: on, 2: off
}  // namespace

// Returns -1 when the product is outside the kernel's envelope (the caller then uses the SIMT kernel), 0 on launch.
int launch_train_gemm_tc(cudaStream_t s, int slices, int M, int N, int K, const float* A, int lda, long long sa, int ta,
                         const float* B, int ldb, long long sb, int tb, float* C, int ldc, long long sc, const float* bias, int k_chunk,
                         int wide_mask) {
    if (N < 16 || K < 16 || M < 32) return -1;
    if (k_chunk > 0 && (k_chunk % KC) != 0) return -1;
    int* err_flag = umma_err_flag();
    ES_CHECK(err_flag, "cannot allocate the device error flag");
    TcGemmParams p;
    p.A = A; p.B = B; p.C = C; p.bias = bias; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.ta = ta; p.tb = tb;
    p.k_chunk = k_chunk; p.sa = sa; p.sb = sb; p.sc = sc; p.wide = wide_mask != 0; p.err = err_flag;
    const int n_tiles = (N + 255) / 256;
    const int per = (N + n_tiles - 1) / n_tiles;
    p.ntp = (per + 15) / 16 * 16;
    const size_t smem = (size_t)STAGES * (2 * GM * KC * 2 + 2 * (size_t)p.ntp * KC * 2);
    static PerDeviceSlot<bool> attr_once; bool& attr_set = attr_once.get();
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(t_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGES * (2 * GM * KC * 2 + 2 * 256 * KC * 2)));
        attr_set = true;
    }
    dim3 grid((M + GM - 1) / GM, n_tiles, slices);          // slices: split-K slices (k_chunk > 0) or batch entries
    t_gemm_tc_kernel<<<grid, 128, smem, s>>>(p);
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace es

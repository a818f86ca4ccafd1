// Generic fused "row GEMM" for the channels-last acoustic path (fp32 SIMT).
//
//   Y[b,t,:] = epilogue( sum_tau  A'[b, t*stride + tau - pad, :] . W_tau )
//
// One CTA owns BM=64 consecutive rows of ONE utterance and one tile of <=256 output
// channels; the operand tile (with its time halo) is staged once in shared memory, the
// weights stream through L1 (every warp of the CTA reads the same W rows), and the whole
// per-row epilogue -- bias, boundary-aware tap bias, activation, scalar head, residual,
// LayerNorm, second residual+LayerNorm, padding mask -- runs in registers with warp-shuffle
// reductions, so each layer touches HBM exactly once for its input and once for its output.
//
// Prologues:  PLAIN   dense conv taps over time (Conv1d k in {1,3,5}, stride 1|2, zero padding)
//             DWCONV  depthwise conv k (groups=C) + bias computed in smem before the GEMM (networks.py:281-282)
//
// This is the reference-precision path for every dense contraction; the tcgen05 kernel in
// es_umma_dec.cu replaces it for the decoder layers (the 85% hot spot).
#include "es_common.cuh"

namespace es {

namespace {

constexpr int BM = 64;          // rows per CTA
constexpr int NTHREADS = 256;   // 8 warps x 8 rows
constexpr int RPW = 8;          // rows per warp
constexpr int kMaxKChunk = 512; // K staged per pass in PLAIN mode

template <int NJ>
__global__ void __launch_bounds__(NTHREADS)
rowgemm_kernel(const RowGemmParams p) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int t0 = blockIdx.x * BM;
    const int col_base = blockIdx.y * (32 * NJ);
    const int K = p.K;
    // PLAIN mode streams K in chunks of KC (one chunk for every layer but the widest projections);
    // DWCONV stages the whole K at once (their K is the decoder width, <= 512).
    const int KC = (p.mode == ROW_PLAIN && K > kMaxKChunk) ? kMaxKChunk : K;
    const int K4 = KC >> 2;
    const int lds = KC + 4;
    float* As = smem;

    // ------------------------------------------------------------------ prologue
    if (p.mode == ROW_PLAIN) {
        // staged per K chunk inside the main loop
    } else {  // ROW_DWCONV
        float* Xs = smem + BM * lds;
        const int half = p.dw_k >> 1;
        const int rows_in = BM + p.dw_k - 1;
        const int first = t0 - half;
        const float* Ab = p.A + (size_t)b * p.n_in * p.lda;
        for (int idx = tid; idx < rows_in * K4; idx += NTHREADS) {
            const int r = idx / K4, c4 = idx - r * K4;
            const int t_in = first + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t_in >= 0 && t_in < p.n_in)
                v = __ldg(reinterpret_cast<const float4*>(Ab + (size_t)t_in * p.lda) + c4);
            *reinterpret_cast<float4*>(Xs + r * lds + c4 * 4) = v;
        }
        __syncthreads();
        for (int idx = tid; idx < BM * K; idx += NTHREADS) {
            const int r = idx / K, c = idx - r * K;
            float acc = __ldg(p.dw_b + c);
            for (int tau = 0; tau < p.dw_k; ++tau)
                acc = fmaf(__ldg(p.dw_w + tau * K + c), Xs[(r + tau) * lds + c], acc);
            As[r * lds + c] = acc;
        }
    }

    // ------------------------------------------------------------------ main loop
    float acc[RPW][NJ];
#pragma unroll
    for (int r = 0; r < RPW; ++r)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[r][j] = 0.f;

    bool jok[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) jok[j] = (col_base + 32 * j) < p.ldw;

    const int r0 = warp * RPW;
    const int taps = (p.mode == ROW_PLAIN) ? p.taps : 1;
    const int stride = (p.mode == ROW_PLAIN) ? p.stride : 1;
    for (int kc0 = 0; kc0 < K; kc0 += KC) {
    if (p.mode == ROW_PLAIN) {
        if (kc0 > 0) __syncthreads();              // previous chunk fully consumed
        const int rows_in = (BM - 1) * p.stride + p.taps;
        const int first = t0 * p.stride - p.pad;
        const float* Ab = p.A + (size_t)b * p.n_in * p.lda + kc0;
        for (int idx = tid; idx < rows_in * K4; idx += NTHREADS) {
            const int r = idx / K4, c4 = idx - r * K4;
            const int t_in = first + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t_in >= 0 && t_in < p.n_in)
                v = __ldg(reinterpret_cast<const float4*>(Ab + (size_t)t_in * p.lda) + c4);
            *reinterpret_cast<float4*>(As + r * lds + c4 * 4) = v;
        }
    }
    __syncthreads();
    for (int tap = 0; tap < taps; ++tap) {
        const float* Wt = p.W + ((size_t)tap * K + kc0) * p.ldw + col_base + lane;
        const float* Arow = As + (r0 * stride + tap) * lds;
        for (int k4 = 0; k4 < KC; k4 += 4) {
            float4 a[RPW];
#pragma unroll
            for (int r = 0; r < RPW; ++r)
                a[r] = *reinterpret_cast<const float4*>(Arow + r * stride * lds + k4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float w[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j)
                    w[j] = jok[j] ? __ldg(Wt + (size_t)(k4 + kk) * p.ldw + 32 * j) : 0.f;
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) acc[r][j] = fmaf(av, w[j], acc[r][j]);
                }
            }
        }
    }
    }  // K chunks

    // ------------------------------------------------------------------ epilogue (per row, in registers)
    float bias[NJ], g1[NJ], b1[NJ], g2[NJ], b2[NJ], dw[NJ];
    bool cok[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int col = col_base + lane + 32 * j;
        cok[j] = col < p.Nout;
        bias[j] = (p.bias && cok[j]) ? __ldg(p.bias + col) : 0.f;
        g1[j] = (p.ln_g && cok[j]) ? __ldg(p.ln_g + col) : 0.f;
        b1[j] = (p.ln_g && cok[j]) ? __ldg(p.ln_b + col) : 0.f;
        g2[j] = (p.ln2_g && cok[j]) ? __ldg(p.ln2_g + col) : 0.f;
        b2[j] = (p.ln2_g && cok[j]) ? __ldg(p.ln2_b + col) : 0.f;
        dw[j] = (p.dot_w && cok[j]) ? __ldg(p.dot_w + col) : 0.f;
    }
    const float inv_n = 1.f / (float)p.Nout;
    const int zero_from = p.zero_from ? p.zero_from[b] : 0x7fffffff;

#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const int t = t0 + r0 + r;
        if (t >= p.n_out) break;                       // warp-uniform
        const size_t row = (size_t)b * p.n_out + t;
        float v[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) v[j] = acc[r][j] + bias[j];
        if (p.tap_bias) {
            for (int tap = 0; tap < taps; ++tap) {
                const int t_in = t * stride + tap - p.pad;
                if (t_in >= 0 && t_in < p.n_in) {
#pragma unroll
                    for (int j = 0; j < NJ; ++j)
                        if (cok[j]) v[j] += __ldg(p.tap_bias + tap * p.ldw + col_base + lane + 32 * j);
                }
            }
        }
        if (p.act1 != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) v[j] = apply_act(v[j], p.act1);
        }
        if (p.dot_out) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) s = fmaf(v[j], dw[j], s);
            s = warp_sum(s) + __ldg(p.dot_b);
            if (p.dot_relu) s = fmaxf(s, 0.f);
            if (lane == 0) p.dot_out[row] = s;
        }
        if (p.res1) {
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                if (cok[j]) v[j] += __ldg(p.res1 + row * p.ldr1 + col_base + lane + 32 * j);
        }
        if (p.ln_g) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) s += cok[j] ? v[j] : 0.f;
            const float mean = warp_sum(s) * inv_n;
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) { const float d = cok[j] ? v[j] - mean : 0.f; q = fmaf(d, d, q); }
            const float rstd = 1.f / sqrtf(warp_sum(q) * inv_n + kLnEps);
#pragma unroll
            for (int j = 0; j < NJ; ++j) v[j] = (v[j] - mean) * rstd * g1[j] + b1[j];
        }
        if (p.act2 != ACT_NONE) {
#pragma unroll
            for (int j = 0; j < NJ; ++j) v[j] = apply_act(v[j], p.act2);
        }
        if (p.res2) {
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                if (cok[j]) v[j] += __ldg(p.res2 + row * p.ldr2 + col_base + lane + 32 * j);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) s += cok[j] ? v[j] : 0.f;
            const float mean = warp_sum(s) * inv_n;
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) { const float d = cok[j] ? v[j] - mean : 0.f; q = fmaf(d, d, q); }
            const float rstd = 1.f / sqrtf(warp_sum(q) * inv_n + kLnEps);
#pragma unroll
            for (int j = 0; j < NJ; ++j) v[j] = (v[j] - mean) * rstd * g2[j] + b2[j];
        }
        const bool zero = (p.row_mask && p.row_mask[row]) || (t >= zero_from);
        if (p.Y) {
#pragma unroll
            for (int j = 0; j < NJ; ++j)
                if (cok[j]) p.Y[row * p.ldy + col_base + lane + 32 * j] = zero ? 0.f : v[j];
        }
    }
}

template <int NJ>
int launch_nj(const RowGemmParams& p, size_t smem, dim3 grid, cudaStream_t stream) {
    static PerDeviceSlot<bool> attr_once; bool& attr_set = attr_once.get();   // function attributes are per device
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(rowgemm_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    rowgemm_kernel<NJ><<<grid, NTHREADS, smem, stream>>>(p);
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace

int launch_rowgemm(const RowGemmParams& p, cudaStream_t stream) {
    ES_CHECK(p.B > 0 && p.n_out > 0 && p.n_in > 0, "empty problem");
    if (p.Nout <= 96 || (p.Nout <= 384 && p.K <= 64)) {     // narrow layers: one thread per row
        const int rc = launch_rowgemm_narrow(p, stream);
        if (rc >= 0) return rc;
    }
    ES_CHECK(p.K % 4 == 0 && p.lda % 4 == 0, "K and lda must be multiples of 4");
    ES_CHECK(p.ldw % 32 == 0 && p.ldw >= p.Nout, "ldw must be Nout padded to a multiple of 32");
    const int slots = p.ldw / 32;
    int NJ = slots <= 1 ? 1 : slots == 2 ? 2 : slots == 3 ? 3 : slots == 4 ? 4 : slots <= 6 ? 6 : 8;
    const int col_tiles = (slots + NJ - 1) / NJ;
    const bool needs_full_row = p.ln_g || p.res2 || p.dot_out;
    ES_CHECK(!needs_full_row || col_tiles == 1, "LayerNorm / scalar head need Nout <= 256");
    ES_CHECK(!(p.res2 && !p.ln2_g), "res2 requires ln2");
    const int KC = (p.mode == ROW_PLAIN && p.K > kMaxKChunk) ? kMaxKChunk : p.K;
    ES_CHECK(p.K % KC == 0, "K must be a multiple of 512 when larger than 512");
    const int lds = KC + 4;
    size_t smem;
    if (p.mode == ROW_PLAIN) {
        ES_CHECK(p.taps >= 1 && p.taps <= ES_MAX_TAPS && (p.stride == 1 || p.stride == 2), "bad taps/stride");
        smem = (size_t)((BM - 1) * p.stride + p.taps) * lds * sizeof(float);
    } else {
        ES_CHECK(p.dw_w && p.dw_b && p.dw_k >= 1 && p.dw_k <= ES_MAX_TAPS && (p.dw_k & 1), "bad depthwise kernel");
        ES_CHECK(p.n_in == p.n_out, "depthwise prologue keeps the length");
        smem = (size_t)(2 * BM + p.dw_k - 1) * lds * sizeof(float);
    }
    ES_CHECK(smem <= 200 * 1024, "operand tile does not fit in shared memory (K too large)");
    dim3 grid((p.n_out + BM - 1) / BM, col_tiles, p.B);
    switch (NJ) {
        case 1: return launch_nj<1>(p, smem, grid, stream);
        case 2: return launch_nj<2>(p, smem, grid, stream);
        case 3: return launch_nj<3>(p, smem, grid, stream);
        case 4: return launch_nj<4>(p, smem, grid, stream);
        case 6: return launch_nj<6>(p, smem, grid, stream);
        default: return launch_nj<8>(p, smem, grid, stream);
    }
}

}  // namespace es

// =============================================================================================
// Narrow-output variant: FOUR THREADS PER ROW (a "quad"), NT/4 consecutive output channels each.
// For the phoneme-side layers of the tiny model (C = 32 / 64) a warp-per-8-rows mapping spends
// ~100 instructions of per-row scalar work (statistics, predicates, addressing) to produce 32
// outputs, and a thread-per-row mapping leaves the GPU with ~1.7 warps per scheduler.  Here a CTA
// of 256 threads owns 64 rows: the input rows sit in shared memory at an odd stride (a quad
// reads one address -> broadcast; the 8 rows of a warp hit 8 different banks), the weights sit in
// shared memory and are read as float4 broadcasts, row statistics need two xor-shuffles.
// PLAIN prologue only (conv taps / stride).
// =============================================================================================
namespace es {
namespace {

constexpr int NB_ROWS = 64;    // rows per CTA
constexpr int NB_THR = 256;    // 4 threads per row

__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

struct RowGemmBatch {
    RowGemmParams p[3];       // independent problems of identical geometry (the three predictors)
};

template <int NT>
__global__ void __launch_bounds__(NB_THR, 2)
rowgemm_narrow_kernel(const RowGemmBatch batch) {
    constexpr int NQ = NT / 4;                         // channels per thread (8, 16 or 24)
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, r = tid >> 2, q = tid & 3;
    const int which = blockIdx.z / batch.p[0].B;
    const RowGemmParams& p = batch.p[which];
    const int b = blockIdx.z - which * batch.p[0].B, t0 = blockIdx.x * NB_ROWS, n0 = blockIdx.y * NT;
    const int K = p.K, lds = K + 1;
    const int rows_in = (NB_ROWS - 1) * p.stride + p.taps;
    float* As = smem;                                  // [rows_in][K+1]
    float* Ws = smem + ((rows_in * lds + 3) & ~3);     // [taps*K][NT]

    {   // stage the input rows (zero outside the sequence) and this column tile of the weights.
        // Loads are issued 8 deep per thread before the first store: the staging costs one global
        // round trip, not one per iteration.
        const int K4 = K >> 2;
        const int first = t0 * p.stride - p.pad;
        const float* Ab = p.A + (size_t)b * p.n_in * p.lda;
        const int totalA = rows_in * K4;
        for (int base = 0; base < totalA; base += NB_THR * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * NB_THR + tid;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < totalA) {
                    const int rr = idx / K4, c4 = idx - rr * K4;
                    const int t_in = first + rr;
                    if (t_in >= 0 && t_in < p.n_in) v[u] = __ldg(reinterpret_cast<const float4*>(Ab + (size_t)t_in * p.lda) + c4);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * NB_THR + tid;
                if (idx < totalA) {
                    const int rr = idx / K4, c4 = idx - rr * K4;
                    float* d = As + rr * lds + c4 * 4;
                    d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
                }
            }
        }
        constexpr int nt4 = NT >> 2;
        const int totalW = p.taps * K * nt4;
        for (int base = 0; base < totalW; base += NB_THR * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * NB_THR + tid;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx < totalW) {
                    const int rk = idx / nt4, c4 = idx - rk * nt4;
                    if (n0 + c4 * 4 < p.ldw) v[u] = __ldg(reinterpret_cast<const float4*>(p.W + (size_t)rk * p.ldw + n0) + c4);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = base + u * NB_THR + tid;
                if (idx < totalW) reinterpret_cast<float4*>(Ws)[idx] = v[u];
            }
        }
    }
    __syncthreads();

    const int t = t0 + r;
    const bool live = t < p.n_out;                     // dead quads keep running: shuffles below are warp-wide
    float acc[NQ];
#pragma unroll
    for (int n = 0; n < NQ; ++n) acc[n] = 0.f;
    for (int tap = 0; tap < p.taps; ++tap) {
        const float* arow = As + (r * p.stride + tap) * lds;
        const float4* wt = reinterpret_cast<const float4*>(Ws + (size_t)tap * K * NT + q * NQ);
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            const float a = arow[k];
#pragma unroll
            for (int n4 = 0; n4 < NQ / 4; ++n4) {
                const float4 w = wt[k * (NT / 4) + n4];
                acc[4 * n4] = fmaf(a, w.x, acc[4 * n4]);
                acc[4 * n4 + 1] = fmaf(a, w.y, acc[4 * n4 + 1]);
                acc[4 * n4 + 2] = fmaf(a, w.z, acc[4 * n4 + 2]);
                acc[4 * n4 + 3] = fmaf(a, w.w, acc[4 * n4 + 3]);
            }
        }
    }
    // ---- epilogue: this thread owns columns c0 .. c0+NQ-1 of row t
    const size_t row = (size_t)b * p.n_out + (live ? t : 0);
    const int c0 = n0 + q * NQ;
    bool cok[NQ];
#pragma unroll
    for (int n = 0; n < NQ; ++n) cok[n] = (c0 + n) < p.Nout;
    if (p.bias) {
#pragma unroll
        for (int n = 0; n < NQ; ++n) if (cok[n]) acc[n] += __ldg(p.bias + c0 + n);
    }
    if (p.tap_bias) {
        for (int tap = 0; tap < p.taps; ++tap) {
            const int t_in = t * p.stride + tap - p.pad;
            if (t_in >= 0 && t_in < p.n_in) {
#pragma unroll
                for (int n = 0; n < NQ; ++n) if (cok[n]) acc[n] += __ldg(p.tap_bias + tap * p.ldw + c0 + n);
            }
        }
    }
    if (p.act1 != ACT_NONE) {
#pragma unroll
        for (int n = 0; n < NQ; ++n) acc[n] = apply_act(acc[n], p.act1);
    }
    if (p.dot_out) {
        float s = 0.f;
#pragma unroll
        for (int n = 0; n < NQ; ++n) if (cok[n]) s = fmaf(acc[n], __ldg(p.dot_w + c0 + n), s);
        s = quad_sum(s) + __ldg(p.dot_b);
        if (live && q == 0) p.dot_out[row] = p.dot_relu ? fmaxf(s, 0.f) : s;
    }
    if (p.res1) {
        const float4* r4 = reinterpret_cast<const float4*>(p.res1 + row * p.ldr1 + c0);
#pragma unroll
        for (int n4 = 0; n4 < NQ / 4; ++n4) {
            if (cok[4 * n4]) {
                const float4 v = __ldg(r4 + n4);
                acc[4 * n4] += v.x; acc[4 * n4 + 1] += v.y; acc[4 * n4 + 2] += v.z; acc[4 * n4 + 3] += v.w;
            }
        }
    }
    if (p.ln_g) {                                        // host guarantees a single column tile here
        const float inv_n = 1.f / (float)p.Nout;
        float s = 0.f;
#pragma unroll
        for (int n = 0; n < NQ; ++n) s += cok[n] ? acc[n] : 0.f;
        const float mean = quad_sum(s) * inv_n;
        float qq = 0.f;
#pragma unroll
        for (int n = 0; n < NQ; ++n) { const float d = cok[n] ? acc[n] - mean : 0.f; qq = fmaf(d, d, qq); }
        const float rstd = 1.f / sqrtf(quad_sum(qq) * inv_n + kLnEps);
#pragma unroll
        for (int n = 0; n < NQ; ++n)
            if (cok[n]) acc[n] = (acc[n] - mean) * rstd * __ldg(p.ln_g + c0 + n) + __ldg(p.ln_b + c0 + n);
    }
    if (p.act2 != ACT_NONE) {
#pragma unroll
        for (int n = 0; n < NQ; ++n) acc[n] = apply_act(acc[n], p.act2);
    }
    if (p.Y && live) {
        const bool zero = (p.row_mask && p.row_mask[row]) || (p.zero_from && t >= p.zero_from[b]);
        float4* y4 = reinterpret_cast<float4*>(p.Y + row * p.ldy + c0);
#pragma unroll
        for (int n4 = 0; n4 < NQ / 4; ++n4) {
            if (cok[4 * n4])
                y4[n4] = zero ? make_float4(0.f, 0.f, 0.f, 0.f)
                              : make_float4(acc[4 * n4], acc[4 * n4 + 1], acc[4 * n4 + 2], acc[4 * n4 + 3]);
        }
    }
}

template <int NT>
int launch_narrow_nt(const RowGemmBatch& batch, size_t smem, dim3 grid, cudaStream_t stream) {
    static PerDeviceSlot<bool> attr_once; bool& attr_set = attr_once.get();   // function attributes are per device
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(rowgemm_narrow_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    rowgemm_narrow_kernel<NT><<<grid, NB_THR, smem, stream>>>(batch);
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace

// Returns -1 if the narrow kernel does not apply (caller falls back to launch_rowgemm).
// `count` (1..3) problems of identical geometry run in ONE launch (grid.z = count * B).
int launch_rowgemm_narrow_batch(const RowGemmParams* ps, int count, cudaStream_t stream) {
    const RowGemmParams& p = ps[0];
    if (count < 1 || count > 3) return -1;
    if (p.mode != ROW_PLAIN || p.res2 || p.Nout % 4 || p.ldy % 4 || (p.res1 && p.ldr1 % 4) || p.K % 4) return -1;
    for (int i = 1; i < count; ++i) {
        const RowGemmParams& o = ps[i];
        if (o.mode != p.mode || o.B != p.B || o.n_in != p.n_in || o.n_out != p.n_out || o.K != p.K || o.Nout != p.Nout ||
            o.ldw != p.ldw || o.taps != p.taps || o.stride != p.stride || o.pad != p.pad || o.res2) return -1;
    }
    int NT;
    if (p.Nout <= 32) NT = 32;
    else if (p.Nout <= 64) NT = 64;
    else if (p.Nout <= 96) NT = 96;
    else if (!(p.ln_g || p.dot_out) && p.Nout % 96 == 0) NT = 96;      // e.g. qkv of block 1: 384 = 4 x 96
    else if (!(p.ln_g || p.dot_out) && p.Nout % 64 == 0) NT = 64;
    else return -1;
    if (p.Nout > NT && (p.Nout % NT)) return -1;
    const int rows_in = (NB_ROWS - 1) * p.stride + p.taps;
    const size_t smem = ((size_t)((rows_in * (p.K + 1) + 3) & ~3) + (size_t)p.taps * p.K * NT) * sizeof(float);
    if (smem > 200 * 1024 || (long long)count * p.B > 65535) return -1;
    RowGemmBatch batch;
    for (int i = 0; i < 3; ++i) batch.p[i] = ps[i < count ? i : 0];
    dim3 grid((p.n_out + NB_ROWS - 1) / NB_ROWS, (p.Nout + NT - 1) / NT, count * p.B);
    switch (NT) {
        case 32: return launch_narrow_nt<32>(batch, smem, grid, stream);
        case 64: return launch_narrow_nt<64>(batch, smem, grid, stream);
        default: return launch_narrow_nt<96>(batch, smem, grid, stream);
    }
}

int launch_rowgemm_narrow(const RowGemmParams& p, cudaStream_t stream) {
    return launch_rowgemm_narrow_batch(&p, 1, stream);
}

}  // namespace es

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
//             GATHER  the length regulator: row t <- fused4[b, upper_bound(cum[b], t)]   (networks.py:228-258)
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
    // GATHER / DWCONV stage the whole K at once (their K is the decoder width, <= 512).
    const int KC = (p.mode == ROW_PLAIN && K > kMaxKChunk) ? kMaxKChunk : K;
    const int K4 = KC >> 2;
    const int lds = KC + 4;
    float* As = smem;

    // ------------------------------------------------------------------ prologue
    if (p.mode == ROW_PLAIN) {
        // staged per K chunk inside the main loop
    } else if (p.mode == ROW_GATHER) {
        int* srcs = reinterpret_cast<int*>(smem + BM * lds);
        if (tid < BM) {
            const int t = t0 + tid;
            int s = -1;
            if (t < p.n_out && t < p.valid_len[b]) {
                // upper_bound: first n with cum[b,n] > t   (FeatureUpsampler == repeat_interleave)
                const int* c = p.cum + (size_t)b * p.n_in;
                int lo = 0, hi = p.n_in;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (__ldg(c + mid) > t) hi = mid; else lo = mid + 1;
                }
                s = lo < p.n_in ? lo : -1;
            }
            srcs[tid] = s;
        }
        __syncthreads();
        const float* Ab = p.A + (size_t)b * p.n_in * p.lda;
        for (int idx = tid; idx < BM * K4; idx += NTHREADS) {
            const int r = idx / K4, c4 = idx - r * K4;
            const int s = srcs[r];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s >= 0) v = __ldg(reinterpret_cast<const float4*>(Ab + (size_t)s * p.lda) + c4);
            *reinterpret_cast<float4*>(As + r * lds + c4 * 4) = v;
        }
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
    static bool attr_set = false;   // per-instantiation; idempotent, so a benign race at worst
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
    } else if (p.mode == ROW_GATHER) {
        ES_CHECK(p.cum && p.valid_len, "gather needs cum and valid_len");
        smem = (size_t)BM * lds * sizeof(float) + BM * sizeof(int);
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

// Differentiable primitives for the training step (BASELINE configs[4]).
//
// The inference path fuses a whole layer -- or the whole phoneme side -- into one tcgen05 kernel and keeps nothing; a
// training step needs the intermediate activations and the adjoint of every operator.  This file is the FIRST CORRECT
// VERSION of that second path: a small set of generic fp32 SIMT kernels, each with its backward, from which
// efficientspeech_b200/train_ops.py composes Phoneme2Mel.forward(train=True) as torch.autograd.Functions (torch keeps
// the tape; every byte of arithmetic is here).  Nothing in it is tuned: tiled GEMM without tensor cores, convolutions as
// im2col + GEMM, one launch per operator (DESIGN.md section 13 states the measured step time next to the FLOP floor).
//
//   es_t_gemm        C[b] (+)= op(A[b]) op(B[b]) (+ bias): Linear / 1x1 conv / attention products and all their adjoints
//   es_t_im2col / es_t_col2im   dense Conv1d (k taps, stride 1 | 2) and ConvTranspose1d as GEMMs; each other's adjoint
//   es_t_dwconv_fwd / _bwd_x / _bwd_w   depthwise Conv1d over time (MelDecoder, networks.py:281)
//   es_t_layernorm_fwd / _bwd  saves the normalised rows and 1/sigma; column sums for the affine gradients
//   es_t_act_fwd / _bwd        ReLU | exact-erf GELU | tanh
//   es_t_softmax_fwd / _bwd    rows of (x * scale)
//   es_t_gather_rows / es_t_scatter_add_rows   embeddings; es_t_expand_rows / es_t_reduce_rows: the length regulator and
//                              its adjoint (a segmented sum over each phoneme's frames: deterministic, no atomics)
//   es_t_colsum      bias / LayerNorm-affine gradients; es_t_axpby, es_t_mask_rows, es_t_copy2d: glue
#include "es_common.cuh"
#include "es_kernels.cuh"

namespace es {
namespace {

constexpr int TB = 256;

// ---------------------------------------------------------------------------------------------------------------- GEMM
constexpr int GT = 64, GK = 16;
__global__ void __launch_bounds__(256)
t_gemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, long long sa, int ta,
              const float* __restrict__ B, int ldb, long long sb, int tb, float* __restrict__ C, int ldc, long long sc,
              const float* __restrict__ bias, int accumulate, int k_chunk) {
    __shared__ float As[GK][GT + 4], Bs[GK][GT + 4];
    const int bz = blockIdx.z;
    if (k_chunk > 0) {          // split-K: slice bz owns k in [bz k_chunk, min(K, (bz + 1) k_chunk)) and writes partial bz
        const long long k_lo = (long long)bz * k_chunk;
        A += ta ? k_lo * lda : k_lo;
        B += tb ? k_lo : k_lo * ldb;
        K = (int)min((long long)k_chunk, (long long)K - k_lo);
    } else {
        A += (long long)bz * sa; B += (long long)bz * sb;
    }
    C += (long long)bz * sc;
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += GK) {
        for (int idx = threadIdx.x; idx < GT * GK; idx += 256) {
            // A(m, k) = ta ? A[k * lda + m] : A[m * lda + k]; pick the contiguous index as the fast one
            int m, k;
            if (ta) { m = idx % GT; k = idx / GT; } else { k = idx % GK; m = idx / GK; }
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < M && gk < K) ? __ldg(ta ? A + (long long)gk * lda + gm : A + (long long)gm * lda + gk) : 0.f;
        }
        for (int idx = threadIdx.x; idx < GT * GK; idx += 256) {
            // B(k, n) = tb ? B[n * ldb + k] : B[k * ldb + n]
            int n, k;
            if (tb) { k = idx % GK; n = idx / GK; } else { n = idx % GT; k = idx / GT; }
            const int gn = n0 + n, gk = k0 + k;
            Bs[k][n] = (gn < N && gk < K) ? __ldg(tb ? B + (long long)gn * ldb + gk : B + (long long)gk * ldb + gn) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j] + (bias ? __ldg(bias + gn) : 0.f);
            float* c = C + (long long)gm * ldc + gn;
            *c = accumulate ? *c + v : v;
        }
    }
}

// ------------------------------------------------------------------------------------------------- im2col / col2im
// cols[(b n_out + t) (C k) + c k + tau] = X[b, t s + tau - p, c] (0 outside): a row is in the order of torch's flattened
// Conv1d weight [Cout][Cin k], so the weight and its gradient never need a permuted copy
__global__ void __launch_bounds__(TB)
t_im2col_kernel(const float* __restrict__ X, float* __restrict__ cols, int B, int n_in, int n_out, int C, int k, int s, int p) {
    const long long total = (long long)B * n_out * k * C;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < total; i += (long long)gridDim.x * TB) {
        const int tau = (int)(i % k);
        const int c = (int)((i / k) % C);
        const long long row = i / ((long long)C * k);
        const int t = (int)(row % n_out), b = (int)(row / n_out);
        const int ti = t * s + tau - p;
        cols[i] = (ti >= 0 && ti < n_in) ? __ldg(X + ((long long)b * n_in + ti) * C + c) : 0.f;
    }
}
// X[b, i, c] (+)= sum over (t, tau) with t s + tau - p == i of cols[b, t, tau, c]   (the adjoint of im2col)
__global__ void __launch_bounds__(TB)
t_col2im_kernel(const float* __restrict__ cols, float* __restrict__ X, int B, int n_in, int n_out, int C, int k, int s, int p,
                int accumulate) {
    const long long total = (long long)B * n_in * C;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < total; i += (long long)gridDim.x * TB) {
        const int c = (int)(i % C);
        const long long row = i / C;
        const int ti = (int)(row % n_in), b = (int)(row / n_in);
        float v = 0.f;
        for (int tau = 0; tau < k; ++tau) {
            const int num = ti + p - tau;
            if (num < 0 || num % s) continue;
            const int t = num / s;
            if (t < n_out) v += __ldg(cols + (((long long)b * n_out + t) * C + c) * k + tau);
        }
        X[i] = accumulate ? X[i] + v : v;
    }
}

// ----------------------------------------------------------------------------------------------------- depthwise conv
// 4 consecutive floats; `aligned` (uniform per launch) selects the 16-byte access.  Parameters live at arbitrary element
// offsets of the optimiser's flat buffer, and activations may be views, so alignment is a run-time property.
__device__ __forceinline__ float4 ld4(const float* p, bool aligned) {
    if (aligned) return __ldg(reinterpret_cast<const float4*>(p));
    return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}
__device__ __forceinline__ void st4(float* p, float4 v, bool aligned) {
    if (aligned) { *reinterpret_cast<float4*>(p) = v; return; }
    p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
}
__device__ __forceinline__ bool aligned16(const void* a, const void* b) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}
// one thread = 4 channels of one frame (C % 4 == 0), 32-bit index arithmetic
__global__ void __launch_bounds__(TB)
t_dwconv_fwd_kernel(const float* __restrict__ X, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ Y,
                    int B, int T, int C, int k) {      // w [C][k] (torch [C,1,k])
    const int C4 = C >> 2, total = B * T * C4, p = k / 2;
    const bool al = aligned16(X, Y);
    for (int i = blockIdx.x * TB + threadIdx.x; i < total; i += gridDim.x * TB) {
        const int c = (i % C4) * 4, row = i / C4, t = row % T;
        float4 v = ld4(bias + c, false);
        for (int tau = 0; tau < k; ++tau) {
            const int ti = t + tau - p;
            if (ti < 0 || ti >= T) continue;
            const float4 x = ld4(X + (size_t)(row + tau - p) * C + c, al);
            v.x = fmaf(__ldg(w + c * k + tau), x.x, v.x);
            v.y = fmaf(__ldg(w + (c + 1) * k + tau), x.y, v.y);
            v.z = fmaf(__ldg(w + (c + 2) * k + tau), x.z, v.z);
            v.w = fmaf(__ldg(w + (c + 3) * k + tau), x.w, v.w);
        }
        st4(Y + (size_t)row * C + c, v, al);
    }
}
__global__ void __launch_bounds__(TB)
t_dwconv_bwd_x_kernel(const float* __restrict__ dY, const float* __restrict__ w, float* __restrict__ dX, int B, int T, int C, int k) {
    const int C4 = C >> 2, total = B * T * C4, p = k / 2;
    const bool al = aligned16(dY, dX);
    for (int i = blockIdx.x * TB + threadIdx.x; i < total; i += gridDim.x * TB) {
        const int c = (i % C4) * 4, row = i / C4, ti = row % T;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int tau = 0; tau < k; ++tau) {
            const int t = ti - tau + p;                 // y[t] used x[t + tau - p]
            if (t < 0 || t >= T) continue;
            const float4 g = ld4(dY + (size_t)(row + p - tau) * C + c, al);
            v.x = fmaf(__ldg(w + c * k + tau), g.x, v.x);
            v.y = fmaf(__ldg(w + (c + 1) * k + tau), g.y, v.y);
            v.z = fmaf(__ldg(w + (c + 2) * k + tau), g.z, v.z);
            v.w = fmaf(__ldg(w + (c + 3) * k + tau), g.w, v.w);
        }
        st4(dX + (size_t)row * C + c, v, al);
    }
}
// ---- sliding-window forms (k known at compile time): every input row is loaded ONCE ---------------------------------
__device__ __forceinline__ float4 fma4(float4 w, float4 x, float4 a) {
    return make_float4(fmaf(w.x, x.x, a.x), fmaf(w.y, x.y, a.y), fmaf(w.z, x.z, a.z), fmaf(w.w, x.w, a.w));
}
// Y[b,t,c] = bias[c] + sum_tau w[c][flip ? K-1-tau : tau] X[b,t+tau-p,c].  One thread = 4 channels x RB consecutive frames
// of one utterance: RB + K - 1 row loads for RB outputs instead of RB K.  flip (with bias == nullptr) is the input
// gradient: dX = conv(dY, reversed taps).
template <int K, int RB>
__global__ void __launch_bounds__(TB)
t_dwconv_win_kernel(const float* __restrict__ X, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ Y,
                    int B, int T, int C, int flip) {
    constexpr int P = K / 2;
    const int C4 = C >> 2, tiles = (T + RB - 1) / RB, total = B * tiles * C4;
    const bool al = aligned16(X, Y);
    for (int i = blockIdx.x * TB + threadIdx.x; i < total; i += gridDim.x * TB) {
        const int c = (i % C4) * 4, bt = i / C4, b = bt / tiles, t0 = (bt - b * tiles) * RB;
        float4 wt[K];
#pragma unroll
        for (int tau = 0; tau < K; ++tau) {
            const int s = flip ? K - 1 - tau : tau;
            wt[tau] = make_float4(__ldg(w + c * K + s), __ldg(w + (c + 1) * K + s), __ldg(w + (c + 2) * K + s), __ldg(w + (c + 3) * K + s));
        }
        float4 acc[RB];
        const float4 b4 = bias ? ld4(bias + c, false) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int o = 0; o < RB; ++o) acc[o] = b4;
        const float* xb = X + (size_t)b * T * C + c;
#pragma unroll
        for (int j = 0; j < RB + K - 1; ++j) {
            const int t = t0 + j - P;
            if (t < 0 || t >= T) continue;
            const float4 x = ld4(xb + (size_t)t * C, al);
#pragma unroll
            for (int o = 0; o < RB; ++o) {
                const int tau = j - o;                  // output t0 + o reads row t0 + o + tau - P
                if (tau >= 0 && tau < K) acc[o] = fma4(wt[tau], x, acc[o]);
            }
        }
        float* yb = Y + (size_t)b * T * C + c;
#pragma unroll
        for (int o = 0; o < RB; ++o)
            if (t0 + o < T) st4(yb + (size_t)(t0 + o) * C, acc[o], al);
    }
}
// Weight / bias gradient partials.  grid (ceil(C/128), slices); block = 32 lanes x 4 channels, 8 row lanes; a row lane
// walks its contiguous share of the slice frame by frame with the K input rows of the current frame in registers.
template <int K>
__global__ void __launch_bounds__(256)
t_dwconv_bwd_w_win_kernel(const float* __restrict__ dY, const float* __restrict__ X, float* __restrict__ part, int B, int T, int C,
                          int rows_per_slice) {
    constexpr int P = K / 2;
    __shared__ float4 red[8][32];
    const int cl = threadIdx.x & 31, c = (blockIdx.x * 32 + cl) * 4, lane_r = threadIdx.x >> 5;
    const int rows = B * T;
    const int s0 = blockIdx.y * rows_per_slice, s1 = min(rows, s0 + rows_per_slice);
    const int per = (s1 - s0 + 7) / 8;
    const int r0 = s0 + lane_r * per, r1 = min(s1, r0 + per);
    const bool al = aligned16(dY, X);
    float4 acc[K + 1], win[K];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i <= K; ++i) acc[i] = z;
    if (c < C && r0 < r1) {
        int t = r0 % T;
        // win[tau] = X[row + tau - P] (zero outside the utterance); filled at the start and whenever an utterance begins
        bool refill = true;
        for (int row = r0; row < r1; ++row) {
            if (refill) {
#pragma unroll
                for (int tau = 0; tau < K - 1; ++tau) {
                    const int ti = t + tau - P;
                    win[tau + 1] = (ti >= 0 && ti < T) ? ld4(X + (size_t)(row + tau - P) * C + c, al) : z;
                }
                refill = false;
            }
#pragma unroll
            for (int tau = 0; tau < K - 1; ++tau) win[tau] = win[tau + 1];
            win[K - 1] = (t + P < T) ? ld4(X + (size_t)(row + P) * C + c, al) : z;
            const float4 g = ld4(dY + (size_t)row * C + c, al);
            acc[K].x += g.x; acc[K].y += g.y; acc[K].z += g.z; acc[K].w += g.w;
#pragma unroll
            for (int tau = 0; tau < K; ++tau) acc[tau] = fma4(g, win[tau], acc[tau]);
            if (++t == T) { t = 0; refill = true; }
        }
    }
    // block reduction over the 8 row lanes, one tap at a time, in a fixed order
    for (int i = 0; i <= K; ++i) {
        red[lane_r][cl] = acc[i];
        __syncthreads();
        if (lane_r == 0 && c < C) {
            float4 tsum = z;
            for (int r = 0; r < 8; ++r) { const float4 v = red[r][cl]; tsum.x += v.x; tsum.y += v.y; tsum.z += v.z; tsum.w += v.w; }
            float* out = part + ((size_t)blockIdx.y * C + c) * (K + 1) + i;
            out[0] = tsum.x; out[K + 1] = tsum.y; out[2 * (K + 1)] = tsum.z; out[3 * (K + 1)] = tsum.w;
        }
        __syncthreads();
    }
}

// dw[c][tau] = sum_{b,t} dY[b,t,c] X[b,t+tau-p,c]; db[c] = sum dY.  grid (ceil(C/32), S row slices); block = 32 channels x 8
// row lanes; every thread keeps the k taps + the bias sum in registers and the block writes ONE partial row
// part[slice][c (k+1) + tau] -- the slices are added afterwards in a fixed order (t_colsum_kernel), so the result does not
// depend on scheduling.
constexpr int DW_KMAX = 7;
__global__ void __launch_bounds__(256)
t_dwconv_bwd_w_kernel(const float* __restrict__ dY, const float* __restrict__ X, float* __restrict__ part, int B, int T, int C, int k,
                      long long rows_per_slice) {
    __shared__ float red[8][32][DW_KMAX + 2];
    const int cl = threadIdx.x & 31, c = blockIdx.x * 32 + cl, lane_r = threadIdx.x >> 5, p = k / 2;
    const long long rows = (long long)B * T;
    const long long r0 = blockIdx.y * rows_per_slice, r1 = min(rows, r0 + rows_per_slice);
    float acc[DW_KMAX + 1];
#pragma unroll
    for (int i = 0; i <= DW_KMAX; ++i) acc[i] = 0.f;
    if (c < C) {
        for (long long row = r0 + lane_r; row < r1; row += 8) {
            const int t = (int)(row % T);
            const float g = __ldg(dY + row * C + c);
            acc[DW_KMAX] += g;
#pragma unroll
            for (int tau = 0; tau < DW_KMAX; ++tau) {
                const int ti = t + tau - p;
                if (tau < k && ti >= 0 && ti < T) acc[tau] = fmaf(g, __ldg(X + (row + tau - p) * C + c), acc[tau]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i <= DW_KMAX; ++i) red[lane_r][cl][i] = acc[i];
    __syncthreads();
    if (lane_r == 0 && c < C) {
        float* out = part + ((long long)blockIdx.y * C + c) * (k + 1);
        for (int i = 0; i <= k; ++i) {
            const int src = i == k ? DW_KMAX : i;
            float t = 0.f;
            for (int r = 0; r < 8; ++r) t += red[r][cl][src];
            out[i] = t;
        }
    }
}
// part summed over slices [C][k+1] -> dw [C][k], db [C]
__global__ void __launch_bounds__(TB)
t_dwconv_split_kernel(const float* __restrict__ sum, float* __restrict__ dw, float* __restrict__ db, int C, int k) {
    const int i = blockIdx.x * TB + threadIdx.x;
    if (i >= C * (k + 1)) return;
    const int c = i / (k + 1), tau = i % (k + 1);
    if (tau == k) db[c] = sum[i]; else dw[c * k + tau] = sum[i];
}

// ---------------------------------------------------------------------------------------------------------- LayerNorm
// one warp per row: xhat = (x - mean) rstd (saved), y = xhat g + b
__global__ void __launch_bounds__(TB)
t_layernorm_fwd_kernel(const float* __restrict__ X, const float* __restrict__ g, const float* __restrict__ be, float* __restrict__ Y,
                       float* __restrict__ xhat, float* __restrict__ rstd, long long rows, int C) {
    const int lane = threadIdx.x & 31;
    for (long long r = (long long)blockIdx.x * (TB / 32) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (TB / 32)) {
        const float* x = X + r * C;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += x[c];
        const float mean = warp_sum(s) / C;
        float q = 0.f;
        for (int c = lane; c < C; c += 32) { const float d = x[c] - mean; q = fmaf(d, d, q); }
        const float rs = rsqrtf(warp_sum(q) / C + kLnEps);
        if (lane == 0) rstd[r] = rs;
        for (int c = lane; c < C; c += 32) {
            const float h = (x[c] - mean) * rs;
            xhat[r * C + c] = h;
            Y[r * C + c] = fmaf(h, __ldg(g + c), __ldg(be + c));
        }
    }
}
// dx = rstd (gdy - mean(gdy) - xhat mean(gdy xhat)), gdy = dy g.  The same pass accumulates this block's share of
// dg[c] = sum_r dy xhat and db[c] = sum_r dy in registers (lane owns c = lane + 32 j) and writes one partial row
// part[block][0|1][C]; the blocks are added afterwards in a fixed order.  NJ = ceil(C / 32) <= 8.
template <int NJ>
__global__ void __launch_bounds__(TB)
t_layernorm_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ xhat, const float* __restrict__ rstd,
                       const float* __restrict__ g, float* __restrict__ dX, float* __restrict__ part, long long rows, int C) {
    __shared__ float red[TB / 32][2][NJ * 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float gg[NJ], sg[NJ], sb[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) { const int c = lane + 32 * j; gg[j] = c < C ? __ldg(g + c) : 0.f; sg[j] = 0.f; sb[j] = 0.f; }
    for (long long r = (long long)blockIdx.x * (TB / 32) + w; r < rows; r += (long long)gridDim.x * (TB / 32)) {
        float dy[NJ], xh[NJ], a = 0.f, b = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int c = lane + 32 * j;
            dy[j] = c < C ? dY[r * C + c] : 0.f;
            xh[j] = c < C ? xhat[r * C + c] : 0.f;
            const float gd = dy[j] * gg[j];
            a += gd;
            b = fmaf(gd, xh[j], b);
            sg[j] = fmaf(dy[j], xh[j], sg[j]);
            sb[j] += dy[j];
        }
        a = warp_sum(a) / C; b = warp_sum(b) / C;
        const float rs = rstd[r];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int c = lane + 32 * j;
            if (c < C) dX[r * C + c] = rs * (dy[j] * gg[j] - a - xh[j] * b);
        }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) { red[w][0][lane + 32 * j] = sg[j]; red[w][1][lane + 32 * j] = sb[j]; }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += TB) {
        const int which = i / C, c = i - which * C;
        float t = 0.f;
        for (int ww = 0; ww < TB / 32; ++ww) t += red[ww][which][c];
        part[((long long)blockIdx.x * 2 + which) * C + c] = t;
    }
}
// out[slice][c] = sum over the slice's rows of A[r,c] (* B[r,c]);  grid (ceil(C/32), slices); block (32 columns, 8 row lanes).
// One slice: the column sums themselves.  Many slices: partial rows, added by a second one-slice launch (fixed order).
__global__ void __launch_bounds__(256)
t_colsum_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ out, long long rows, int C, int accumulate,
                long long rows_per_slice, int ld) {
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + (threadIdx.x & 31), lane_r = threadIdx.x >> 5;
    const long long r0 = blockIdx.y * rows_per_slice, r1 = min(rows, r0 + rows_per_slice);
    float s = 0.f;
    if (c < C)
        for (long long r = r0 + lane_r; r < r1; r += 8) s += Bm ? A[r * ld + c] * Bm[r * ld + c] : A[r * ld + c];
    part[lane_r][threadIdx.x & 31] = s;
    __syncthreads();
    if (lane_r == 0 && c < C) {
        float t = 0.f;
        for (int r = 0; r < 8; ++r) t += part[r][threadIdx.x & 31];
        float* o = out + (long long)blockIdx.y * C + c;
        *o = accumulate ? *o + t : t;
    }
}

// -------------------------------------------------------------------------------------------------------- activations
__device__ __forceinline__ float act_f(float x, int kind) {
    if (kind == ACT_RELU) return fmaxf(x, 0.f);
    if (kind == ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
    if (kind == ACT_TANH) return tanhf(x);
    return x;
}
__global__ void __launch_bounds__(TB)
t_act_fwd_kernel(const float* __restrict__ X, float* __restrict__ Y, long long n, int kind) {
    const long long n4 = ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) & 15) ? 0 : n >> 2;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < n4; i += (long long)gridDim.x * TB) {
        float4 v = reinterpret_cast<const float4*>(X)[i];
        v.x = act_f(v.x, kind); v.y = act_f(v.y, kind); v.z = act_f(v.z, kind); v.w = act_f(v.w, kind);
        reinterpret_cast<float4*>(Y)[i] = v;
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * TB + threadIdx.x; i < n; i += (long long)gridDim.x * TB) Y[i] = act_f(X[i], kind);
}
__device__ __forceinline__ float act_df(float v, int kind) {
    if (kind == ACT_RELU) return v > 0.f ? 1.f : 0.f;
    if (kind == ACT_TANH) return 1.f - v * v;
    if (kind == ACT_GELU) return 0.5f * (1.f + erff(v * 0.70710678118654752440f)) + v * 0.3989422804014327f * expf(-0.5f * v * v);
    return 1.f;
}
// ReLU / tanh use the saved OUTPUT y, GELU the saved INPUT x
__global__ void __launch_bounds__(TB)
t_act_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ saved, float* __restrict__ dX, long long n, int kind) {
    const long long n4 = ((reinterpret_cast<uintptr_t>(dY) | reinterpret_cast<uintptr_t>(saved) | reinterpret_cast<uintptr_t>(dX)) & 15) ? 0 : n >> 2;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < n4; i += (long long)gridDim.x * TB) {
        const float4 g = reinterpret_cast<const float4*>(dY)[i], v = reinterpret_cast<const float4*>(saved)[i];
        reinterpret_cast<float4*>(dX)[i] = make_float4(g.x * act_df(v.x, kind), g.y * act_df(v.y, kind), g.z * act_df(v.z, kind),
                                                       g.w * act_df(v.w, kind));
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * TB + threadIdx.x; i < n; i += (long long)gridDim.x * TB)
        dX[i] = dY[i] * act_df(saved[i], kind);
}

// ------------------------------------------------------------------------------------------------------------ softmax
__global__ void __launch_bounds__(TB)
t_softmax_fwd_kernel(const float* __restrict__ X, float* __restrict__ Y, long long rows, int n, float scale) {
    const int lane = threadIdx.x & 31;
    for (long long r = (long long)blockIdx.x * (TB / 32) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (TB / 32)) {
        float m = -INFINITY;
        for (int c = lane; c < n; c += 32) m = fmaxf(m, X[r * n + c] * scale);
        m = warp_max(m);
        float s = 0.f;
        for (int c = lane; c < n; c += 32) s += expf(X[r * n + c] * scale - m);
        s = warp_sum(s);
        for (int c = lane; c < n; c += 32) Y[r * n + c] = expf(X[r * n + c] * scale - m) / s;
    }
}
__global__ void __launch_bounds__(TB)
t_softmax_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ Y, float* __restrict__ dX, long long rows, int n, float scale) {
    const int lane = threadIdx.x & 31;
    for (long long r = (long long)blockIdx.x * (TB / 32) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (TB / 32)) {
        float d = 0.f;
        for (int c = lane; c < n; c += 32) d = fmaf(dY[r * n + c], Y[r * n + c], d);
        d = warp_sum(d);
        for (int c = lane; c < n; c += 32) dX[r * n + c] = scale * Y[r * n + c] * (dY[r * n + c] - d);
    }
}

// ------------------------------------------------------------------------------------------------- gathers / scatters
__global__ void __launch_bounds__(TB)
t_gather_rows_kernel(const float* __restrict__ table, const int32_t* __restrict__ idx, float* __restrict__ out, long long rows, int C) {
    const long long total = rows * C;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < total; i += (long long)gridDim.x * TB) {
        const long long r = i / C;
        const int j = __ldg(idx + r);
        out[i] = j >= 0 ? __ldg(table + (long long)j * C + (i - r * C)) : 0.f;
    }
}
__global__ void __launch_bounds__(TB)
t_scatter_add_rows_kernel(const float* __restrict__ dOut, const int32_t* __restrict__ idx, float* __restrict__ dTable, long long rows,
                          int C, int skip_index) {
    const long long total = rows * C;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < total; i += (long long)gridDim.x * TB) {
        const long long r = i / C;
        const int j = __ldg(idx + r);
        if (j >= 0 && j != skip_index) atomicAdd(dTable + (long long)j * C + (i - r * C), dOut[i]);
    }
}
// length regulator: out[b, t, :] = in[b, n(t), :] for t < cum[b, N-1] (n(t): first n with cum[b,n] > t), 0 beyond
__global__ void __launch_bounds__(TB)
t_expand_rows_kernel(const float* __restrict__ in, const int32_t* __restrict__ cum, float* __restrict__ out, int B, int N, int T, int C) {
    const long long total = (long long)B * T * C;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < total; i += (long long)gridDim.x * TB) {
        const int c = (int)(i % C);
        const long long row = i / C;
        const int t = (int)(row % T), b = (int)(row / T);
        const int32_t* cb = cum + (long long)b * N;
        int lo = 0, hi = N;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(cb + mid) > t) hi = mid; else lo = mid + 1; }
        out[i] = lo < N ? __ldg(in + ((long long)b * N + lo) * C + c) : 0.f;
    }
}
// its adjoint: dIn[b, n, :] = sum of dOut[b, t, :] over the frames of phoneme n
__global__ void __launch_bounds__(TB)
t_reduce_rows_kernel(const float* __restrict__ dOut, const int32_t* __restrict__ cum, float* __restrict__ dIn, int B, int N, int T, int C) {
    const long long total = (long long)B * N * C;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < total; i += (long long)gridDim.x * TB) {
        const int c = (int)(i % C);
        const long long row = i / C;
        const int n = (int)(row % N), b = (int)(row / N);
        const int t0 = n ? __ldg(cum + row - 1) : 0;
        const int t1 = min(__ldg(cum + row), T);
        float s = 0.f;
        for (int t = t0; t < t1; ++t) s += __ldg(dOut + ((long long)b * T + t) * C + c);
        dIn[i] = s;
    }
}

// --------------------------------------------------------------------------------------------------------------- glue
__global__ void __launch_bounds__(TB)
t_axpby_kernel(const float* __restrict__ X, const float* __restrict__ Yin, float* __restrict__ out, long long n, float a, float b) {
    const long long n4 = (!Yin || ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Yin) | reinterpret_cast<uintptr_t>(out)) & 15)) ? 0 : n >> 2;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < n4; i += (long long)gridDim.x * TB) {
        const float4 x = reinterpret_cast<const float4*>(X)[i], y = reinterpret_cast<const float4*>(Yin)[i];
        reinterpret_cast<float4*>(out)[i] = make_float4(a * x.x + b * y.x, a * x.y + b * y.y, a * x.z + b * y.z, a * x.w + b * y.w);
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * TB + threadIdx.x; i < n; i += (long long)gridDim.x * TB)
        out[i] = a * X[i] + (Yin ? b * Yin[i] : 0.f);
}
__global__ void __launch_bounds__(TB)
t_mask_rows_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask, float* __restrict__ Y, long long rows, int C) {
    const long long total = rows * C;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < total; i += (long long)gridDim.x * TB) Y[i] = mask[i / C] ? 0.f : X[i];
}
__global__ void __launch_bounds__(TB)
t_copy2d_kernel(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, long long rows, int cols, int accumulate) {
    const long long total = rows * cols;
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < total; i += (long long)gridDim.x * TB) {
        const long long r = i / cols;
        const int c = (int)(i - r * cols);
        float* d = dst + r * ldd + c;
        *d = accumulate ? *d + src[r * lds + c] : src[r * lds + c];
    }
}

// torch.bucketize(v, bins) (right=False): the number of boundaries < v, i.e. the first i with bins[i] >= v
__global__ void __launch_bounds__(TB)
t_bucketize_kernel(const float* __restrict__ v, const float* __restrict__ bins, int n_bins, int32_t* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * TB + threadIdx.x; i < n; i += (long long)gridDim.x * TB) {
        const float x = v[i];
        int lo = 0, hi = n_bins;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(bins + mid) >= x) hi = mid; else lo = mid + 1; }
        out[i] = lo;
    }
}

inline unsigned blocks_for(long long n) {
    long long b = (n + TB - 1) / TB;
    return (unsigned)(b < 1 ? 1 : (b > 148 * 32 ? 148 * 32 : b));
}

// slices of a long column reduction: ~512 rows each, at most 256
inline int slices_for(long long rows) {
    if (rows <= 2048) return 1;
    long long s = (rows + 511) / 512;
    return (int)(s > 256 ? 256 : s);
}
inline int colsum(cudaStream_t st, const float* A, const float* Bm, float* out, long long rows, int C, int accumulate, float* ws,
                  size_t ws_floats) {
    const int S = slices_for(rows);
    if (S == 1) {
        t_colsum_kernel<<<dim3((C + 31) / 32, 1), 256, 0, st>>>(A, Bm, out, rows, C, accumulate, rows, C);
        return cudaGetLastError() == cudaSuccess ? 0 : 1;
    }
    if (!ws || ws_floats < (size_t)S * C) return 2;
    const long long per = (rows + S - 1) / S;
    t_colsum_kernel<<<dim3((C + 31) / 32, S), 256, 0, st>>>(A, Bm, ws, rows, C, 0, per, C);
    t_colsum_kernel<<<dim3((C + 31) / 32, 1), 256, 0, st>>>(ws, nullptr, out, S, C, accumulate, S, C);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
inline int ln_bwd_blocks(long long rows) {
    long long b = (rows + TB / 32 - 1) / (TB / 32);
    return (int)(b < 1 ? 1 : (b > 148 * 4 ? 148 * 4 : b));
}

}  // namespace
}  // namespace es

using namespace es;
#define ST static_cast<cudaStream_t>(stream)

namespace { std::atomic<int> g_train_tc{1}; }

extern "C" {

void es_t_set_tensor_core(int enable) { g_train_tc.store(enable ? 1 : 0, std::memory_order_relaxed); }

int es_t_gemm(void* stream, int batch, int M, int N, int K, const float* A, int lda, long long stride_a, int trans_a,
              const float* B, int ldb, long long stride_b, int trans_b, float* C, int ldc, long long stride_c,
              const float* bias, int accumulate, int k_chunk, int grad_mask) {
    ES_CHECK(A && B && C && batch >= 1 && M >= 1 && N >= 1 && K >= 1 && batch <= 65535 && k_chunk >= 0, "bad arguments");
    ES_CHECK(k_chunk == 0 || (long long)(batch - 1) * k_chunk < K, "split-K: empty slice");
    if (g_train_tc.load(std::memory_order_relaxed) && !accumulate) {
        const int rc = launch_train_gemm_tc(ST, batch, M, N, K, A, lda, stride_a, trans_a, B, ldb, stride_b, trans_b, C, ldc, stride_c, bias,
                                            k_chunk, grad_mask);
        if (rc >= 0) return rc;
    }
    dim3 grid((N + GT - 1) / GT, (M + GT - 1) / GT, batch);
    ES_CHECK(grid.y <= 65535, "M too large for one launch");
    t_gemm_kernel<<<grid, 256, 0, ST>>>(M, N, K, A, lda, stride_a, trans_a, B, ldb, stride_b, trans_b, C, ldc, stride_c, bias, accumulate,
                                        k_chunk);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_im2col(void* stream, const float* X, float* cols, int B, int n_in, int n_out, int C, int k, int s, int p) {
    ES_CHECK(X && cols && s >= 1 && k >= 1, "bad arguments");
    t_im2col_kernel<<<blocks_for((long long)B * n_out * k * C), TB, 0, ST>>>(X, cols, B, n_in, n_out, C, k, s, p);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_col2im(void* stream, const float* cols, float* X, int B, int n_in, int n_out, int C, int k, int s, int p, int accumulate) {
    ES_CHECK(X && cols && s >= 1 && k >= 1, "bad arguments");
    t_col2im_kernel<<<blocks_for((long long)B * n_in * C), TB, 0, ST>>>(cols, X, B, n_in, n_out, C, k, s, p, accumulate);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_dwconv_fwd(void* stream, const float* X, const float* w, const float* bias, float* Y, int B, int T, int C, int k) {
    ES_CHECK(X && w && bias && Y, "null tensor");
    ES_CHECK(C % 4 == 0 && (long long)B * T * C < (1ll << 31), "depthwise conv: C must be a multiple of 4, B T C below 2^31");
    constexpr int RB = 8;
    const long long work = (long long)B * ((T + RB - 1) / RB) * (C / 4);
    if (k == 3)      t_dwconv_win_kernel<3, RB><<<blocks_for(work), TB, 0, ST>>>(X, w, bias, Y, B, T, C, 0);
    else if (k == 5) t_dwconv_win_kernel<5, RB><<<blocks_for(work), TB, 0, ST>>>(X, w, bias, Y, B, T, C, 0);
    else if (k == 7) t_dwconv_win_kernel<7, RB><<<blocks_for(work), TB, 0, ST>>>(X, w, bias, Y, B, T, C, 0);
    else             t_dwconv_fwd_kernel<<<blocks_for((long long)B * T * C / 4), TB, 0, ST>>>(X, w, bias, Y, B, T, C, k);
    ES_LAUNCH_OK();
    return 0;
}
// slices of the depthwise weight-gradient reduction: 256 frames each (8 row lanes x 32), at most 1024
static int dw_slices(long long rows) {
    long long sl = (rows + 255) / 256;
    return (int)(sl < 1 ? 1 : (sl > 1024 ? 1024 : sl));
}
size_t es_t_dwconv_bwd_workspace_floats(int B, int T, int C, int k) {
    return (size_t)(dw_slices((long long)B * T) + 1) * C * (k + 1);
}
int es_t_dwconv_bwd(void* stream, const float* dY, const float* X, const float* w, float* dX, float* dw, float* db, int B, int T, int C, int k,
                    float* ws, size_t ws_floats) {
    ES_CHECK(dY && X && w && dX && dw && db && ws, "null tensor");
    ES_CHECK(k >= 1 && k <= DW_KMAX, "depthwise kernel size above 7");
    ES_CHECK(ws_floats >= es_t_dwconv_bwd_workspace_floats(B, T, C, k), "workspace too small");
    ES_CHECK(C % 4 == 0 && (long long)B * T * C < (1ll << 31), "depthwise conv: C must be a multiple of 4, B T C below 2^31");
    constexpr int RB = 8;
    const long long work = (long long)B * ((T + RB - 1) / RB) * (C / 4);
    if (k == 3)      t_dwconv_win_kernel<3, RB><<<blocks_for(work), TB, 0, ST>>>(dY, w, nullptr, dX, B, T, C, 1);
    else if (k == 5) t_dwconv_win_kernel<5, RB><<<blocks_for(work), TB, 0, ST>>>(dY, w, nullptr, dX, B, T, C, 1);
    else if (k == 7) t_dwconv_win_kernel<7, RB><<<blocks_for(work), TB, 0, ST>>>(dY, w, nullptr, dX, B, T, C, 1);
    else             t_dwconv_bwd_x_kernel<<<blocks_for((long long)B * T * C / 4), TB, 0, ST>>>(dY, w, dX, B, T, C, k);
    ES_LAUNCH_OK();
    const long long rows = (long long)B * T;
    const int S = dw_slices(rows), W = C * (k + 1);
    const int per_slice = (int)((rows + S - 1) / S);
    float* sum = ws + (size_t)S * W;
    const dim3 gw((C / 4 + 31) / 32, S);
    if (k == 3)      t_dwconv_bwd_w_win_kernel<3><<<gw, 256, 0, ST>>>(dY, X, ws, B, T, C, per_slice);
    else if (k == 5) t_dwconv_bwd_w_win_kernel<5><<<gw, 256, 0, ST>>>(dY, X, ws, B, T, C, per_slice);
    else if (k == 7) t_dwconv_bwd_w_win_kernel<7><<<gw, 256, 0, ST>>>(dY, X, ws, B, T, C, per_slice);
    else             t_dwconv_bwd_w_kernel<<<dim3((C + 31) / 32, S), 256, 0, ST>>>(dY, X, ws, B, T, C, k, per_slice);
    ES_LAUNCH_OK();
    t_colsum_kernel<<<dim3((W + 31) / 32, 1), 256, 0, ST>>>(ws, nullptr, sum, S, W, 0, S, W);
    ES_LAUNCH_OK();
    t_dwconv_split_kernel<<<(W + TB - 1) / TB, TB, 0, ST>>>(sum, dw, db, C, k);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_layernorm_fwd(void* stream, const float* X, const float* g, const float* b, float* Y, float* xhat, float* rstd, long long rows, int C) {
    ES_CHECK(X && g && b && Y && xhat && rstd, "null tensor");
    t_layernorm_fwd_kernel<<<blocks_for(rows * 32), TB, 0, ST>>>(X, g, b, Y, xhat, rstd, rows, C);
    ES_LAUNCH_OK();
    return 0;
}
size_t es_t_layernorm_bwd_workspace_floats(long long rows, int C) { return (size_t)ln_bwd_blocks(rows) * 2 * C; }
int es_t_layernorm_bwd(void* stream, const float* dY, const float* xhat, const float* rstd, const float* g, float* dX, float* dg, float* db,
                       long long rows, int C, float* ws, size_t ws_floats) {
    ES_CHECK(dY && xhat && rstd && g && dX && dg && db && ws, "null tensor");
    ES_CHECK(C >= 1 && C <= 256, "LayerNorm width above 256");
    ES_CHECK(ws_floats >= es_t_layernorm_bwd_workspace_floats(rows, C), "workspace too small");
    const int G = ln_bwd_blocks(rows);
    if (C <= 32)       t_layernorm_bwd_kernel<1><<<G, TB, 0, ST>>>(dY, xhat, rstd, g, dX, ws, rows, C);
    else if (C <= 64)  t_layernorm_bwd_kernel<2><<<G, TB, 0, ST>>>(dY, xhat, rstd, g, dX, ws, rows, C);
    else if (C <= 128) t_layernorm_bwd_kernel<4><<<G, TB, 0, ST>>>(dY, xhat, rstd, g, dX, ws, rows, C);
    else               t_layernorm_bwd_kernel<8><<<G, TB, 0, ST>>>(dY, xhat, rstd, g, dX, ws, rows, C);
    ES_LAUNCH_OK();
    // ws rows are [block][dg | db][C]: two reductions over G rows with a row stride of 2C
    t_colsum_kernel<<<dim3((C + 31) / 32, 1), 256, 0, ST>>>(ws, nullptr, dg, G, C, 0, G, 2 * C);
    ES_LAUNCH_OK();
    t_colsum_kernel<<<dim3((C + 31) / 32, 1), 256, 0, ST>>>(ws + C, nullptr, db, G, C, 0, G, 2 * C);
    ES_LAUNCH_OK();
    return 0;
}
size_t es_t_colsum_workspace_floats(long long rows, int C) { return slices_for(rows) > 1 ? (size_t)slices_for(rows) * C : 0; }
int es_t_colsum(void* stream, const float* A, const float* B, float* out, long long rows, int C, int accumulate, float* ws, size_t ws_floats) {
    ES_CHECK(A && out, "null tensor");
    const int rc = colsum(ST, A, B, out, rows, C, accumulate, ws, ws_floats);
    ES_CHECK(rc != 2, "workspace too small");
    ES_CHECK(rc == 0, "launch failed");
    return 0;
}
int es_t_act_fwd(void* stream, const float* X, float* Y, long long n, int kind) {
    ES_CHECK(X && Y, "null tensor");
    t_act_fwd_kernel<<<blocks_for(n), TB, 0, ST>>>(X, Y, n, kind);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_act_bwd(void* stream, const float* dY, const float* saved, float* dX, long long n, int kind) {
    ES_CHECK(dY && saved && dX, "null tensor");
    t_act_bwd_kernel<<<blocks_for(n), TB, 0, ST>>>(dY, saved, dX, n, kind);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_softmax_fwd(void* stream, const float* X, float* Y, long long rows, int n, float scale) {
    ES_CHECK(X && Y, "null tensor");
    t_softmax_fwd_kernel<<<blocks_for(rows * 32), TB, 0, ST>>>(X, Y, rows, n, scale);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_softmax_bwd(void* stream, const float* dY, const float* Y, float* dX, long long rows, int n, float scale) {
    ES_CHECK(dY && Y && dX, "null tensor");
    t_softmax_bwd_kernel<<<blocks_for(rows * 32), TB, 0, ST>>>(dY, Y, dX, rows, n, scale);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_gather_rows(void* stream, const float* table, const int32_t* idx, float* out, long long rows, int C) {
    ES_CHECK(table && idx && out, "null tensor");
    t_gather_rows_kernel<<<blocks_for(rows * C), TB, 0, ST>>>(table, idx, out, rows, C);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_scatter_add_rows(void* stream, const float* d_out, const int32_t* idx, float* d_table, long long rows, int C, int skip_index) {
    ES_CHECK(d_out && idx && d_table, "null tensor");
    t_scatter_add_rows_kernel<<<blocks_for(rows * C), TB, 0, ST>>>(d_out, idx, d_table, rows, C, skip_index);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_expand_rows(void* stream, const float* in, const int32_t* cum, float* out, int B, int N, int T, int C) {
    ES_CHECK(in && cum && out, "null tensor");
    t_expand_rows_kernel<<<blocks_for((long long)B * T * C), TB, 0, ST>>>(in, cum, out, B, N, T, C);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_reduce_rows(void* stream, const float* d_out, const int32_t* cum, float* d_in, int B, int N, int T, int C) {
    ES_CHECK(d_out && cum && d_in, "null tensor");
    t_reduce_rows_kernel<<<blocks_for((long long)B * N * C), TB, 0, ST>>>(d_out, cum, d_in, B, N, T, C);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_bucketize(void* stream, const float* v, const float* bins, int n_bins, int32_t* out, long long n) {
    ES_CHECK(v && bins && out, "null tensor");
    t_bucketize_kernel<<<blocks_for(n), TB, 0, ST>>>(v, bins, n_bins, out, n);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_axpby(void* stream, const float* X, const float* Y, float* out, long long n, float a, float b) {
    ES_CHECK(X && out, "null tensor");
    t_axpby_kernel<<<blocks_for(n), TB, 0, ST>>>(X, Y, out, n, a, b);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_mask_rows(void* stream, const float* X, const uint8_t* mask, float* Y, long long rows, int C) {
    ES_CHECK(X && mask && Y, "null tensor");
    t_mask_rows_kernel<<<blocks_for(rows * C), TB, 0, ST>>>(X, mask, Y, rows, C);
    ES_LAUNCH_OK();
    return 0;
}
int es_t_copy2d(void* stream, const float* src, int lds, float* dst, int ldd, long long rows, int cols, int accumulate) {
    ES_CHECK(src && dst, "null tensor");
    t_copy2d_kernel<<<blocks_for(rows * cols), TB, 0, ST>>>(src, lds, dst, ldd, rows, cols, accumulate);
    ES_LAUNCH_OK();
    return 0;
}

}  // extern "C"

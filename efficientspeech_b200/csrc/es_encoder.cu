// Phoneme-side kernels that are not plain row GEMMs: table-fused embedding+merge conv,
// unmasked multi-head attention, the folded Fuse stage, variance embedding + duration
// rounding + integer scan (first half of the length regulator), and the materialising
// gather (second half).  All fp32 / int32, channels-last.
#include "es_common.cuh"
#include "es_kernels.cuh"

namespace es {

// -----------------------------------------------------------------------------------------
// K_EMB: x0[b,t,:] = sum_tau Tab[tau][ id[b, t+tau-pad] ][:]      (zero outside the sequence)
// Tab[tau] = E . (W1x1 . Wk[:,:,tau])^T  is built at pack time, so the embedding lookup
// (networks.py:54), the dense merge conv (:65) and the 1x1 projection (:66) of encoder
// block 0 are k gathers of C floats per phoneme.
// -----------------------------------------------------------------------------------------
__global__ void embed_merge_kernel(const int32_t* __restrict__ ids, const float* __restrict__ tab,
                                   float* __restrict__ out, int B, int N, int C, int k, int n_symbols) {
    const int C4 = C >> 2;
    const long long total = (long long)B * N * C4;
    const int pad = k >> 1;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % C4);
        const long long row = idx / C4;
        const int t = (int)(row % N);
        const long long b = row / N;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int tau = 0; tau < k; ++tau) {
            const int ti = t + tau - pad;
            if (ti < 0 || ti >= N) continue;
            int id = __ldg(ids + b * N + ti);
            id = min(max(id, 0), n_symbols - 1);
            const float4 v = __ldg(reinterpret_cast<const float4*>(tab + ((size_t)tau * n_symbols + id) * C) + c4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(out)[idx] = acc;
    }
}

int launch_embed_merge(const int32_t* ids, const float* tab, float* out, int B, int N, int C, int k,
                       int n_symbols, cudaStream_t s) {
    const long long total = (long long)B * N * (C / 4);
    const int threads = 256;
    long long want = (total + threads - 1) / threads;
    const int blocks = (int)(want < 148 * 16 ? want : 148 * 16);
    embed_merge_kernel<<<blocks, threads, 0, s>>>(ids, tab, out, B, N, C, k, n_symbols);
    ES_LAUNCH_OK();
    return 0;
}

// -----------------------------------------------------------------------------------------
// Pooled padding mask of encoder block 1 (blocks.py:51-57): pad the mask with True to a
// multiple of `pool`, then max over groups of `pool`.
// -----------------------------------------------------------------------------------------
__global__ void pool_mask_kernel(const uint8_t* __restrict__ mask, uint8_t* __restrict__ out,
                                 int B, int N, int n1, int pool) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * n1) return;
    const int b = idx / n1, j = idx - b * n1;
    uint8_t m = 0;
    for (int q = 0; q < pool; ++q) {
        const int t = j * pool + q;
        m |= (t < N) ? mask[(size_t)b * N + t] : (uint8_t)1;
    }
    out[idx] = m ? 1 : 0;
}

int launch_pool_mask(const uint8_t* mask, uint8_t* out, int B, int N, int n1, int pool, cudaStream_t s) {
    const int total = B * n1;
    pool_mask_kernel<<<(total + 255) / 256, 256, 0, s>>>(mask, out, B, N, n1, pool);
    ES_LAUNCH_OK();
    return 0;
}

// -----------------------------------------------------------------------------------------
// K_ATT: O[b, q, h*C:(h+1)*C] = softmax_k( Q_h[q] . K_h[k] * scale ) V_h      (blocks.py:46-65)
// Every head is full width C; the softmax runs over ALL n keys, padding included -- the
// reference builds an attention mask and never applies it (blocks.py:59-63); reproduced.
// One CTA = 32 query rows of one (utterance, head).  Scores for the 32 rows live in shared
// memory ([32][n]); K and V stream through a 32-key tile.
// qkv layout: [B, n, 3*H*C] with channel order [q|k|v][head][c]  (blocks.py:45).
// -----------------------------------------------------------------------------------------
constexpr int ATT_Q = 32;
constexpr int ATT_KT = 32;

__global__ void __launch_bounds__(256)
attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int n, int C, int H, float scale) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * ATT_Q;
    const int ldc = C + 4;                       // padded row stride (floats), keeps float4 alignment
    const int lds = n + 1;                       // score row stride
    float* Qs = smem;                            // [32][ldc]
    float* KVs = Qs + ATT_Q * ldc;               // [32][ldc]
    float* S = KVs + ATT_KT * ldc;               // [32][lds]
    const int C4 = C >> 2;
    const size_t ldq = (size_t)3 * H * C;
    const float* base = qkv + (size_t)b * n * ldq;
    const int qoff = h * C, koff = (H + h) * C, voff = (2 * H + h) * C;

    for (int idx = tid; idx < ATT_Q * C4; idx += 256) {
        const int r = idx / C4, c4 = idx - r * C4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q0 + r < n) v = __ldg(reinterpret_cast<const float4*>(base + (size_t)(q0 + r) * ldq + qoff) + c4);
        *reinterpret_cast<float4*>(Qs + r * ldc + c4 * 4) = v;
    }
    // ---- scores
    const int qi = tid >> 3, kg = tid & 7;       // 32 query rows x 8 key groups (keys kg, kg+8, kg+16, kg+24)
    for (int k0 = 0; k0 < n; k0 += ATT_KT) {
        __syncthreads();
        for (int idx = tid; idx < ATT_KT * C4; idx += 256) {
            const int r = idx / C4, c4 = idx - r * C4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + r < n) v = __ldg(reinterpret_cast<const float4*>(base + (size_t)(k0 + r) * ldq + koff) + c4);
            *reinterpret_cast<float4*>(KVs + r * ldc + c4 * 4) = v;
        }
        __syncthreads();
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        const float* qrow = Qs + qi * ldc;
        for (int c = 0; c < C; c += 4) {
            const float4 q = *reinterpret_cast<const float4*>(qrow + c);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const float4 kv = *reinterpret_cast<const float4*>(KVs + (kg + 8 * m) * ldc + c);
                s[m] = fmaf(q.x, kv.x, s[m]); s[m] = fmaf(q.y, kv.y, s[m]);
                s[m] = fmaf(q.z, kv.z, s[m]); s[m] = fmaf(q.w, kv.w, s[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int key = k0 + kg + 8 * m;
            if (key < n) S[qi * lds + key] = s[m] * scale;
        }
    }
    __syncthreads();
    // ---- softmax over all n keys: one warp handles 4 rows
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int r = warp * 4; r < warp * 4 + 4; ++r) {
            float* srow = S + r * lds;
            float mx = -INFINITY;
            for (int k = lane; k < n; k += 32) mx = fmaxf(mx, srow[k]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int k = lane; k < n; k += 32) { const float e = expf(srow[k] - mx); srow[k] = e; sum += e; }
            sum = warp_sum(sum);
            const float inv = 1.f / sum;
            for (int k = lane; k < n; k += 32) srow[k] *= inv;
        }
    }
    // ---- O = P V : thread (qi, cg) owns columns cg*4 + 32*m
    const int cg = tid & 7;
    float4 o[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) o[m] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nm = C >> 5;                        // column groups of 32 (C is a multiple of 32, <= 256)
    for (int k0 = 0; k0 < n; k0 += ATT_KT) {
        __syncthreads();
        for (int idx = tid; idx < ATT_KT * C4; idx += 256) {
            const int r = idx / C4, c4 = idx - r * C4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + r < n) v = __ldg(reinterpret_cast<const float4*>(base + (size_t)(k0 + r) * ldq + voff) + c4);
            *reinterpret_cast<float4*>(KVs + r * ldc + c4 * 4) = v;
        }
        __syncthreads();
        const int kmax = min(ATT_KT, n - k0);
        for (int k = 0; k < kmax; ++k) {
            const float pw = S[qi * lds + k0 + k];
            const float* vrow = KVs + k * ldc + cg * 4;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                if (m < nm) {
                    const float4 v = *reinterpret_cast<const float4*>(vrow + 32 * m);
                    o[m].x = fmaf(pw, v.x, o[m].x); o[m].y = fmaf(pw, v.y, o[m].y);
                    o[m].z = fmaf(pw, v.z, o[m].z); o[m].w = fmaf(pw, v.w, o[m].w);
                }
            }
        }
    }
    if (q0 + qi < n) {
        float* orow = out + ((size_t)b * n + q0 + qi) * ((size_t)H * C) + h * C + cg * 4;
#pragma unroll
        for (int m = 0; m < 8; ++m)
            if (m < nm) *reinterpret_cast<float4*>(orow + 32 * m) = o[m];
    }
}

int launch_attention(const float* qkv, float* out, int B, int n, int C, int H, float scale, cudaStream_t s) {
    ES_CHECK(C % 32 == 0 && C <= 256, "attention width must be a multiple of 32 and <= 256");
    const size_t smem = ((size_t)(ATT_Q + ATT_KT) * (C + 4) + (size_t)ATT_Q * (n + 1)) * sizeof(float);
    ES_CHECK(smem <= 200 * 1024, "phoneme sequence too long for the attention score tile");
    static PerDeviceSlot<bool> attr_once; bool& attr_set = attr_once.get();   // function attributes are per device
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    dim3 grid((n + ATT_Q - 1) / ATT_Q, H, B);
    attention_kernel<<<grid, 256, smem, s>>>(qkv, out, n, C, H, scale);
    ES_LAUNCH_OK();
    return 0;
}

// -----------------------------------------------------------------------------------------
// K_FUSE (networks.py:189-219), every linear map folded at pack time:
//   fused[t] = c + A0 feat0[t] + sum_{(j,tau): 2j+tau = t, 0 <= j < n1} ( G_tau feat1[j] + g_tau ),  masked.
// One warp per output row; lanes over the d output channels.
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fuse_kernel(const float* __restrict__ f0, const float* __restrict__ f1, const float* __restrict__ a0,
            const float* __restrict__ g, const float* __restrict__ gb, const float* __restrict__ cst,
            const uint8_t* __restrict__ mask, float* __restrict__ out, int B, int N, int n1, int d, int k) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (long long)B * N) return;
    const int t = (int)(row % N);
    const long long b = row / N;
    const int nj = d >> 5;                        // d in {32,64,96,128}
    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = (j < nj) ? __ldg(cst + lane + 32 * j) : 0.f;
    const float* x0 = f0 + row * d;
    for (int kk = 0; kk < d; ++kk) {
        const float a = __ldg(x0 + kk);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nj) acc[j] = fmaf(a, __ldg(a0 + (size_t)kk * d + lane + 32 * j), acc[j]);
    }
    const int d2 = 2 * d;
    for (int tau = (t & 1); tau < k; tau += 2) {  // 2j + tau = t  ->  tau has the parity of t
        const int j1 = (t - tau) >> 1;
        if (t - tau < 0 || j1 >= n1) continue;
        const float* x1 = f1 + ((size_t)b * n1 + j1) * d2;
        const float* gt = g + (size_t)tau * d2 * d;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < nj) acc[j] += __ldg(gb + tau * d + lane + 32 * j);
        for (int kk = 0; kk < d2; ++kk) {
            const float a = __ldg(x1 + kk);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (j < nj) acc[j] = fmaf(a, __ldg(gt + (size_t)kk * d + lane + 32 * j), acc[j]);
        }
    }
    const bool zero = mask && mask[row];
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (j < nj) out[row * d + lane + 32 * j] = zero ? 0.f : acc[j];
}

// Thread-per-row variant for d <= 64: all folded weights in shared memory (warp-wide broadcast
// reads), the d accumulators of a row in registers.
template <int D>
__global__ void __launch_bounds__(128)
fuse_narrow_kernel(const float* __restrict__ f0, const float* __restrict__ f1, const float* __restrict__ a0,
                   const float* __restrict__ g, const float* __restrict__ gb, const float* __restrict__ cst,
                   const uint8_t* __restrict__ mask, float* __restrict__ out, int B, int N, int n1, int k) {
    extern __shared__ __align__(16) float sw[];
    float* sA0 = sw;                       // [D][D]
    float* sG = sA0 + D * D;               // [k][2D][D]
    float* sGb = sG + k * 2 * D * D;       // [k][D]
    float* sC = sGb + k * D;               // [D]
    for (int i = threadIdx.x; i < D * D; i += 128) sA0[i] = __ldg(a0 + i);
    for (int i = threadIdx.x; i < k * 2 * D * D; i += 128) sG[i] = __ldg(g + i);
    for (int i = threadIdx.x; i < k * D; i += 128) sGb[i] = __ldg(gb + i);
    for (int i = threadIdx.x; i < D; i += 128) sC[i] = __ldg(cst + i);
    __syncthreads();
    const long long row = (long long)blockIdx.x * 128 + threadIdx.x;
    if (row >= (long long)B * N) return;
    const int t = (int)(row % N);
    const long long b = row / N;
    float acc[D];
#pragma unroll
    for (int n = 0; n < D; ++n) acc[n] = sC[n];
    {
        const float4* x4 = reinterpret_cast<const float4*>(f0 + row * D);
#pragma unroll 2
        for (int k4 = 0; k4 < D / 4; ++k4) {
            const float4 xv = __ldg(x4 + k4);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4* w4 = reinterpret_cast<const float4*>(sA0 + (k4 * 4 + kk) * D);
#pragma unroll
                for (int n4 = 0; n4 < D / 4; ++n4) {
                    const float4 w = w4[n4];
                    acc[4 * n4] = fmaf(xs[kk], w.x, acc[4 * n4]); acc[4 * n4 + 1] = fmaf(xs[kk], w.y, acc[4 * n4 + 1]);
                    acc[4 * n4 + 2] = fmaf(xs[kk], w.z, acc[4 * n4 + 2]); acc[4 * n4 + 3] = fmaf(xs[kk], w.w, acc[4 * n4 + 3]);
                }
            }
        }
    }
    for (int tau = (t & 1); tau < k; tau += 2) {      // 2j + tau = t  ->  tau has the parity of t
        const int j1 = (t - tau) >> 1;
        if (t - tau < 0 || j1 >= n1) continue;
#pragma unroll
        for (int n = 0; n < D; ++n) acc[n] += sGb[tau * D + n];
        const float4* x4 = reinterpret_cast<const float4*>(f1 + ((size_t)b * n1 + j1) * (2 * D));
        const float* gt = sG + (size_t)tau * 2 * D * D;
#pragma unroll 2
        for (int k4 = 0; k4 < (2 * D) / 4; ++k4) {
            const float4 xv = __ldg(x4 + k4);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const float4* w4 = reinterpret_cast<const float4*>(gt + (k4 * 4 + kk) * D);
#pragma unroll
                for (int n4 = 0; n4 < D / 4; ++n4) {
                    const float4 w = w4[n4];
                    acc[4 * n4] = fmaf(xs[kk], w.x, acc[4 * n4]); acc[4 * n4 + 1] = fmaf(xs[kk], w.y, acc[4 * n4 + 1]);
                    acc[4 * n4 + 2] = fmaf(xs[kk], w.z, acc[4 * n4 + 2]); acc[4 * n4 + 3] = fmaf(xs[kk], w.w, acc[4 * n4 + 3]);
                }
            }
        }
    }
    const bool zero = mask && mask[row];
    float4* o4 = reinterpret_cast<float4*>(out + row * D);
#pragma unroll
    for (int n4 = 0; n4 < D / 4; ++n4)
        o4[n4] = zero ? make_float4(0.f, 0.f, 0.f, 0.f)
                      : make_float4(acc[4 * n4], acc[4 * n4 + 1], acc[4 * n4 + 2], acc[4 * n4 + 3]);
}

int launch_fuse(const float* f0, const float* f1, const float* a0, const float* g, const float* gb,
                const float* cst, const uint8_t* mask, float* out, int B, int N, int n1, int d, int k,
                cudaStream_t s) {
    ES_CHECK(d % 32 == 0 && d <= 128, "fuse width must be 32..128");
    const long long rows = (long long)B * N;
    const size_t smem = ((size_t)d * d + (size_t)k * 2 * d * d + (size_t)k * d + d) * sizeof(float);
    if ((d == 32 || d == 64) && smem <= 200 * 1024) {
        static PerDeviceSlot<bool> attr_once; bool& attr_set = attr_once.get();   // function attributes are per device
        if (!attr_set) {
            ES_CUDA(cudaFuncSetAttribute(fuse_narrow_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            ES_CUDA(cudaFuncSetAttribute(fuse_narrow_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set = true;
        }
        const unsigned grid = (unsigned)((rows + 127) / 128);
        if (d == 32) fuse_narrow_kernel<32><<<grid, 128, smem, s>>>(f0, f1, a0, g, gb, cst, mask, out, B, N, n1, k);
        else fuse_narrow_kernel<64><<<grid, 128, smem, s>>>(f0, f1, a0, g, gb, cst, mask, out, B, N, n1, k);
        ES_LAUNCH_OK();
        return 0;
    }
    fuse_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(f0, f1, a0, g, gb, cst, mask, out, B, N, n1, d, k);
    ES_LAUNCH_OK();
    return 0;
}

// -----------------------------------------------------------------------------------------
// K_VAR: one CTA per utterance.
//   fused4[b,n,:] = [ fused | pitch_table[bucketize(pitch)] | energy_table[bucketize(energy)] | dur_feat ]
//                   (embeddings and dur_feat zeroed on padded phonemes)        networks.py:349-377
//   dur_int = clamp(mask ? 0 : (tgt | rint(dur_pred)), 0, 65535)              networks.py:379-384,234
//   dur_cum = inclusive scan(dur_int) (warp-shuffle scan + carried prefix); mel_len = last
// bucketize(v, bins) with right=False = #{j : bins[j] < v}                     networks.py:130-141
// -----------------------------------------------------------------------------------------
__device__ __forceinline__ int bucketize_left(const float* __restrict__ bins, int nb, float v) {
    int lo = 0, hi = nb;                          // first index with bins[idx] >= v
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(bins + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;                                    // NaN compares false everywhere -> 0; torch gives nb, ids stay in range
}

__global__ void __launch_bounds__(256)
variance_scan_kernel(const float* __restrict__ fused, const float* __restrict__ dur_feat,
                     const float* __restrict__ pitch_pred, const float* __restrict__ energy_pred,
                     const float* __restrict__ dur_pred,
                     const float* __restrict__ pitch_tgt, const float* __restrict__ energy_tgt,
                     const int32_t* __restrict__ dur_tgt, const uint8_t* __restrict__ mask,
                     const float* __restrict__ pbins, const float* __restrict__ ptab,
                     const float* __restrict__ ebins, const float* __restrict__ etab,
                     float* __restrict__ fused4, int32_t* __restrict__ dur_int, int32_t* __restrict__ dur_cum,
                     int32_t* __restrict__ mel_len, int N, int d) {
    extern __shared__ int sidx[];                 // [2][rows of this block] bucket indices
    __shared__ int warp_tot[8];
    __shared__ int carry_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t rb = (size_t)b * N;
    // grid.y splits the concat rows of one utterance (64 blocks of 256 threads left most of the GPU idle at B = 64);
    // the duration scan needs the whole utterance and is done by the y == 0 block alone
    const int rpc = (N + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * rpc, r1 = min(N, r0 + rpc);
    if (blockIdx.y == 0) {
        if (tid == 0) carry_s = 0;
        __syncthreads();
        // ---- durations + scan, 256 phonemes per pass
        for (int n0 = 0; n0 < N; n0 += 256) {
            const int n = n0 + tid;
            int dv = 0;
            if (n < N) {
                const bool pad = mask && mask[rb + n];
                float df = dur_tgt ? (float)dur_tgt[rb + n] : rintf(dur_pred[rb + n]);   // torch.round = half-to-even
                if (pad) df = 0.f;
                df = fminf(fmaxf(df, 0.f), 65535.f);   // clamp(min=0); upper clamp only guards int overflow
                dv = (int)df;                           // .int()
                dur_int[rb + n] = dv;
            }
            int incl = dv;                              // warp-shuffle inclusive scan
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            if (lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            int prefix = carry_s;
            for (int w = 0; w < warp; ++w) prefix += warp_tot[w];
            if (n < N) dur_cum[rb + n] = prefix + incl;
            __syncthreads();
            if (tid == 255) carry_s = prefix + incl;
            __syncthreads();
        }
        if (tid == 0) mel_len[b] = carry_s;
    }
    // ---- bucket indices of this block's rows
    for (int n = r0 + tid; n < r1; n += 256) {
        const float pv = pitch_tgt ? pitch_tgt[rb + n] : pitch_pred[rb + n];
        const float ev = energy_tgt ? energy_tgt[rb + n] : energy_pred[rb + n];
        sidx[n - r0] = bucketize_left(pbins, d - 1, pv);
        sidx[rpc + n - r0] = bucketize_left(ebins, d - 1, ev);
    }
    __syncthreads();
    // ---- concat rows: 4 channels (one float4) per thread step; group boundaries are multiples of d (>= 32)
    const int d4 = 4 * d, q4 = d4 >> 2;
    for (int idx = tid; idx < (r1 - r0) * q4; idx += 256) {
        const int nl = idx / q4, n = r0 + nl, c = (idx - nl * q4) * 4;
        const bool pad = mask && mask[rb + n];
        const int grp = c / d, cc = c - grp * d;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (grp == 0) v = *reinterpret_cast<const float4*>(fused + (rb + n) * d + cc);        // already masked by the fuse stage
        else if (!pad) {
            if (grp == 1) v = __ldg(reinterpret_cast<const float4*>(ptab + (size_t)sidx[nl] * d + cc));
            else if (grp == 2) v = __ldg(reinterpret_cast<const float4*>(etab + (size_t)sidx[rpc + nl] * d + cc));
            else v = *reinterpret_cast<const float4*>(dur_feat + (rb + n) * d + cc);
        }
        *reinterpret_cast<float4*>(fused4 + (rb + n) * d4 + c) = v;
    }
}

int launch_variance_scan(const float* fused, const float* dur_feat, const float* pitch_pred,
                         const float* energy_pred, const float* dur_pred, const float* pitch_tgt,
                         const float* energy_tgt, const int32_t* dur_tgt, const uint8_t* mask,
                         const es_predictor_w_t& pw, const es_predictor_w_t& ew, float* fused4,
                         int32_t* dur_int, int32_t* dur_cum, int32_t* mel_len, int B, int N, int d,
                         cudaStream_t s) {
    // enough row chunks to put ~4 blocks on every SM at this batch size, at least 8 rows each
    int chunks = (4 * 148 + B - 1) / B;
    chunks = chunks < 1 ? 1 : (chunks > (N + 7) / 8 ? (N + 7) / 8 : chunks);
    const int rpc = (N + chunks - 1) / chunks;
    const size_t smem = (size_t)2 * rpc * sizeof(int);
    ES_CHECK(smem <= 40 * 1024, "phoneme sequence too long");
    variance_scan_kernel<<<dim3(B, chunks), 256, smem, s>>>(fused, dur_feat, pitch_pred, energy_pred, dur_pred, pitch_tgt,
                                              energy_tgt, dur_tgt, mask, pw.bins, pw.table, ew.bins, ew.table,
                                              fused4, dur_int, dur_cum, mel_len, N, d);
    ES_LAUNCH_OK();
    return 0;
}

// -----------------------------------------------------------------------------------------
// K_LR (materialising form of FeatureUpsampler, networks.py:228-258):
//   src[b,t] = upper_bound(dur_cum[b,:], t) for t < mel_len[b], else -1
//   features[b,t,:] = fused4[b, src, :] or 0;  frame_mask[b,t] = padding | phoneme_mask[b,src]
// One warp per frame: lane 0 searches, the warp copies the row with 128-bit accesses.
// -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
length_regulate_kernel(const float* __restrict__ fused4, const int32_t* __restrict__ cum,
                       const uint8_t* __restrict__ pmask, float* __restrict__ feats,
                       uint8_t* __restrict__ fmask, int32_t* __restrict__ src_out, int B, int N, int T, int C) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (long long)B * T) return;
    const int t = (int)(row % T);
    const long long b = row / T;
    const int32_t* c = cum + b * N;
    int s = -1;
    if (lane == 0) {
        const int total = __ldg(c + N - 1);
        if (t < total) {
            int lo = 0, hi = N;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (__ldg(c + mid) > t) hi = mid; else lo = mid + 1;
            }
            s = lo;
        }
    }
    s = __shfl_sync(0xffffffffu, s, 0);
    if (lane == 0) {
        if (src_out) src_out[row] = s;
        if (fmask) fmask[row] = (s < 0) ? 1 : (pmask ? pmask[b * N + s] : 0);
    }
    if (feats) {
        float4* dst = reinterpret_cast<float4*>(feats + row * C);
        const int C4 = C >> 2;
        if (s >= 0) {
            const float4* sp = reinterpret_cast<const float4*>(fused4 + (b * N + s) * C);
            for (int i = lane; i < C4; i += 32) dst[i] = __ldg(sp + i);
        } else {
            for (int i = lane; i < C4; i += 32) dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

int launch_length_regulate(const float* fused4, const int32_t* cum, const uint8_t* pmask, float* feats,
                           uint8_t* fmask, int32_t* src, int B, int N, int T, int C, cudaStream_t s) {
    const long long rows = (long long)B * T;
    if (rows == 0) return 0;
    length_regulate_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(fused4, cum, pmask, feats, fmask, src, B, N, T, C);
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace es

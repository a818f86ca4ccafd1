// HiFi-GAN generator (V2 geometry and any other ResBlock1 configuration): mel [B,80,T] -> waveform [B, T * prod(rates)].
// The step AFTER the acoustic path: hifigan/models.py:84-134 (Generator), :18-57
// (ResBlock1), used at model.py:161-162.
//
//   x = conv_pre(mel)                                                   Conv1d(80 -> C0, k 7, pad 3)
//   for each upsampling stage i:  x = ConvTranspose1d(lrelu_0.1(x))     C -> C/2, kernel k_i, stride u_i, pad (k_i-u_i)/2
//                                 x = mean over the n ResBlock1_j(x)    kernels (3, 7, 11), dilations (1, 3, 5)
//        ResBlock1: 3 x [ xt = c1_dilated(lrelu_0.1(x)); xt = c2(lrelu_0.1(xt)); x = xt + x ]
//   wav = tanh(conv_post(lrelu_0.01(x)))                                Conv1d(C -> 1, k 7, pad 3)
//
// Channels-first [B, C, L] fp32 like the reference: the channel counts are small (128 ... 8) and the lengths long, so a
// warp's lanes run along time (coalesced) and every thread register-blocks 4 consecutive samples x 8 output channels.
// One generic kernel does every Conv1d (any odd kernel / dilation) with the leaky ReLU of its input applied while the
// tile is staged, and bias, residual add, the 1/n resblock mean (as an accumulate into the stage sum) or tanh fused
// into the store; a second one does the transposed convolutions in gather form (each output sample takes the
// ceil(k/u) taps of its phase).  Input channels are streamed through shared memory 8 at a time together with their
// weights ([ci][tap][co]: a thread's 8 output channels are two 16-byte broadcast loads).  fp32 FMA throughout: the
// reference's fp32 arithmetic, bit-compatible up to summation order.
//
// (The 64- and 32-channel stages are 80 % of the 29 GFLOP per 768-frame utterance and could run on the tensor cores;
// the 16- and 8-channel stages are too narrow for an MMA tile.  This first version is SIMT everywhere -- DESIGN.md.)
#include <new>

#include "es_common.cuh"
#include "es_umma.cuh"      // packed fp32x2 arithmetic (fma.rn.f32x2): two FMAs per issue slot

namespace es {
namespace {

using umma::f32x2;
using umma::fma2;
using umma::pk2;
using umma::up2;

constexpr int HG_THREADS = 256;
constexpr int PT = 4;            // consecutive samples per thread
constexpr int CO = 8;            // output channels per thread
constexpr int CI_CHUNK = 8;      // input channels staged per round
constexpr int HG_MAX_K = 16;

struct ConvParams {
    const float* x;              // input [B, Cin, Lin] (or, for the mel, any strides: sb / sc / st in floats)
    long long sb, sc, st;
    const float* w;              // [Cout][Cin][K] (Conv1d) or [Cin][Cout][K] (ConvTranspose1d)
    const float* bias;           // [Cout]
    const float* res;            // optional residual [B, Cout, Lout]
    float* y;                    // [B, Cout_store, Lout]
    int Cin, Cout, Cout_store;   // Cout_store <= Cout: channels actually written (conv_post computes a padded group)
    int Lin, Lout, K, dil, pad, stride;
    float slope_in;              // leaky-ReLU slope applied to the input (1 = identity)
    float out_scale;             // y = (acc + bias + res) * out_scale ...
    int accumulate;              // ... added to what y already holds
    int tanh_out;
};

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// Conv1d, stride 1, zero padding.  grid (ceil(Lout / TL), B); thread (tx, ty): samples t0 + 4 tx .. +3, channels 8 ty .. +7.
// KT: compile-time kernel size (3 / 7 / 11: the resblock kernels, taps fully unrolled) or 0 = run-time p.K;
// DIL1: dilation 1 -- the 4 samples x K taps of a thread read K + 3 consecutive staged samples ONCE per input channel
// (a register window) instead of 4 K.
template <int KT, bool DIL1>
__global__ void __launch_bounds__(HG_THREADS)
hg_conv_kernel(const ConvParams p) {
    extern __shared__ __align__(16) float hsm[];
    const int K = KT ? KT : p.K;
    const int dil = DIL1 ? 1 : p.dil;
    const int n_cg = (p.Cout + CO - 1) / CO;
    const int TX = HG_THREADS / n_cg, TL = TX * PT;
    const int in_w = TL + (K - 1) * dil;                      // staged samples per input channel
    const int in_ld = (in_w + 3) & ~3;
    const int ldw = n_cg * CO;                                // 8 .. 128: a power of two, divides 256
    float* xs = hsm;                                          // [CI_CHUNK][in_ld]
    float* ws = hsm + CI_CHUNK * in_ld;                       // [CI_CHUNK][K][ldw]
    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const int b = blockIdx.y, t0 = blockIdx.x * TL;
    const float* xb = p.x + (long long)b * p.sb;
    const int wco = tid & (ldw - 1), wr0 = tid / ldw, wstep = HG_THREADS / ldw;   // weight staging: column, first row, row step

    // accumulators as packed channel pairs: acc2[i][c2] = (sample i, channels 2 c2 and 2 c2 + 1); one FFMA2 = two FMAs
    f32x2 acc2[PT][CO / 2];
#pragma unroll
    for (int i = 0; i < PT; ++i)
#pragma unroll
        for (int c = 0; c < CO / 2; ++c) acc2[i][c] = 0ull;

    for (int c0 = 0; c0 < p.Cin; c0 += CI_CHUNK) {
        // stage 8 input channels (leaky ReLU applied here, zeros outside the signal) and their weights [ci][tap][co]
#pragma unroll 1
        for (int ci = 0; ci < CI_CHUNK; ++ci) {
            const float* xc = xb + (long long)(c0 + ci) * p.sc;
            for (int s = tid; s < in_w; s += HG_THREADS) {
                const int t = t0 - p.pad + s;
                float v = 0.f;
                if (t >= 0 && t < p.Lin) v = lrelu(__ldg(xc + (long long)t * p.st), p.slope_in);
                xs[ci * in_ld + s] = v;
            }
        }
        for (int r = wr0; r < CI_CHUNK * K; r += wstep) {
            const int ci = r / K, j = r - ci * K;
            ws[r * ldw + wco] = wco < p.Cout ? __ldg(p.w + ((long long)wco * p.Cin + c0 + ci) * K + j) : 0.f;
        }
        __syncthreads();
#pragma unroll 1
        for (int ci = 0; ci < CI_CHUNK; ++ci) {
            const float* xr = xs + ci * in_ld + tx * PT;
            const float* wr = ws + ci * K * ldw + ty * CO;
            if (DIL1 && KT) {
                f32x2 win[PT + (KT ? KT : 1) - 1];                     // (x, x): the sample broadcast over a channel pair
#pragma unroll
                for (int k = 0; k < PT + KT - 1; ++k) { const float v = xr[k]; win[k] = pk2(v, v); }
#pragma unroll
                for (int j = 0; j < KT; ++j) {
                    const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(wr + j * ldw);
                    const ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(wr + j * ldw + 4);
                    const f32x2 wv[CO / 2] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
                    for (int i = 0; i < PT; ++i)
#pragma unroll
                        for (int c = 0; c < CO / 2; ++c) acc2[i][c] = fma2(win[i + j], wv[c], acc2[i][c]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < (KT ? KT : 1); ++j) {
                    for (int jj = (KT ? j : 0); jj < (KT ? j + 1 : K); ++jj) {
                        const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(wr + jj * ldw);
                        const ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(wr + jj * ldw + 4);
                        const f32x2 wv[CO / 2] = {w0.x, w0.y, w1.x, w1.y};
                        f32x2 xv[PT];
#pragma unroll
                        for (int i = 0; i < PT; ++i) { const float v = xr[i + jj * dil]; xv[i] = pk2(v, v); }
#pragma unroll
                        for (int i = 0; i < PT; ++i)
#pragma unroll
                            for (int c = 0; c < CO / 2; ++c) acc2[i][c] = fma2(xv[i], wv[c], acc2[i][c]);
                    }
                }
            }
        }
        __syncthreads();
    }
    float acc[PT][CO];
#pragma unroll
    for (int i = 0; i < PT; ++i)
#pragma unroll
        for (int c = 0; c < CO / 2; ++c) {
            const float2 v = up2(acc2[i][c]);
            acc[i][2 * c] = v.x;
            acc[i][2 * c + 1] = v.y;
        }
    const int t = t0 + tx * PT;
#pragma unroll
    for (int c = 0; c < CO; ++c) {
        const int co = ty * CO + c;
        if (co >= p.Cout_store) continue;
        const float bv = __ldg(p.bias + co);
        const long long o = ((long long)b * p.Cout_store + co) * p.Lout + t;
        if (t + PT <= p.Lout && (o & 3) == 0) {
            float4 v = make_float4(acc[0][c] + bv, acc[1][c] + bv, acc[2][c] + bv, acc[3][c] + bv);
            if (p.res) {
                const float4 r = __ldg(reinterpret_cast<const float4*>(p.res + o));
                v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            }
            v.x *= p.out_scale; v.y *= p.out_scale; v.z *= p.out_scale; v.w *= p.out_scale;
            if (p.tanh_out) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
            if (p.accumulate) {
                const float4 y = *reinterpret_cast<const float4*>(p.y + o);
                v.x += y.x; v.y += y.y; v.z += y.z; v.w += y.w;
            }
            *reinterpret_cast<float4*>(p.y + o) = v;
        } else {
#pragma unroll
            for (int i = 0; i < PT; ++i) {
                if (t + i >= p.Lout) break;
                float v = acc[i][c] + bv;
                if (p.res) v += __ldg(p.res + o + i);
                v *= p.out_scale;
                if (p.tanh_out) v = tanhf(v);
                if (p.accumulate) v += p.y[o + i];
                p.y[o + i] = v;
            }
        }
    }
}

// ConvTranspose1d in gather form: y[co, t] = b[co] + sum_ci sum_{j = (t + pad) mod u, += u, < K} lrelu(x[ci, (t + pad - j) / u]) w[ci, co, j]
// thread (tx, ty): the u consecutive output samples t0 + u tx + r (r < u <= 8: one per tap PHASE), channels 8 ty .. +7.
// All lanes of a warp then use the same taps at the same time (weights are broadcast loads, input samples consecutive
// words) -- with one sample per lane the 8 phases of a warp read 8 different weight rows: an 8-way bank conflict per load.
constexpr int UP_MAX = 8;
__global__ void __launch_bounds__(HG_THREADS)
hg_upsample_kernel(const ConvParams p) {
    extern __shared__ __align__(16) float hsm[];
    const int n_cg = (p.Cout + CO - 1) / CO;
    const int TX = HG_THREADS / n_cg;
    const int u = p.stride, TP = TX * u;
    // input offsets a tap can reach: off = (r + pad - j) / u (exact), r < u, j < K
    const int off_lo = -((p.K - 1 - p.pad + u - 1) / u), off_hi = (u - 1 + p.pad) / u;
    const int in_w = TX + off_hi - off_lo;
    const int in_ld = (in_w + 3) & ~3;
    const int ldw = n_cg * CO;
    float* xs = hsm;                                          // [CI_CHUNK][in_ld]
    float* ws = hsm + CI_CHUNK * in_ld;                       // [CI_CHUNK][K][ldw]
    const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
    const int b = blockIdx.y, t0 = blockIdx.x * TP;           // t0 is a multiple of u
    const int s_base = t0 / u + off_lo;                       // input sample staged at index 0
    const float* xb = p.x + (long long)b * p.sb;
    const int wco = tid & (ldw - 1), wr0 = tid / ldw, wstep = HG_THREADS / ldw;

    float acc[UP_MAX][CO];
#pragma unroll
    for (int r = 0; r < UP_MAX; ++r)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[r][c] = 0.f;

    for (int c0 = 0; c0 < p.Cin; c0 += CI_CHUNK) {
#pragma unroll 1
        for (int ci = 0; ci < CI_CHUNK; ++ci) {
            const float* xc = xb + (long long)(c0 + ci) * p.sc;
            for (int k = tid; k < in_w; k += HG_THREADS) {
                const int sidx = s_base + k;
                float v = 0.f;
                if (sidx >= 0 && sidx < p.Lin) v = lrelu(__ldg(xc + (long long)sidx * p.st), p.slope_in);
                xs[ci * in_ld + k] = v;
            }
        }
        for (int r = wr0; r < CI_CHUNK * p.K; r += wstep) {
            const int ci = r / p.K, j = r - ci * p.K;
            ws[r * ldw + wco] = wco < p.Cout ? __ldg(p.w + ((long long)(c0 + ci) * p.Cout + wco) * p.K + j) : 0.f;
        }
        __syncthreads();
#pragma unroll 1
        for (int ci = 0; ci < CI_CHUNK; ++ci) {
            const float* xr = xs + ci * in_ld + tx - off_lo;
            const float* wr = ws + ci * p.K * ldw + ty * CO;
#pragma unroll
            for (int r = 0; r < UP_MAX; ++r) {
                if (r < u) {
                    for (int j = (r + p.pad) % u; j < p.K; j += u) {          // warp-uniform
                        const float xv = xr[(r + p.pad - j) / u];              // exact division; staged zeros outside the signal
                        const float4 w0 = *reinterpret_cast<const float4*>(wr + j * ldw);
                        const float4 w1 = *reinterpret_cast<const float4*>(wr + j * ldw + 4);
                        acc[r][0] = fmaf(xv, w0.x, acc[r][0]); acc[r][1] = fmaf(xv, w0.y, acc[r][1]);
                        acc[r][2] = fmaf(xv, w0.z, acc[r][2]); acc[r][3] = fmaf(xv, w0.w, acc[r][3]);
                        acc[r][4] = fmaf(xv, w1.x, acc[r][4]); acc[r][5] = fmaf(xv, w1.y, acc[r][5]);
                        acc[r][6] = fmaf(xv, w1.z, acc[r][6]); acc[r][7] = fmaf(xv, w1.w, acc[r][7]);
                    }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < CO; ++c) {
        const int co = ty * CO + c;
        if (co >= p.Cout_store) continue;
        const float bv = __ldg(p.bias + co);
        float* yr = p.y + ((long long)b * p.Cout_store + co) * p.Lout + t0 + u * tx;
#pragma unroll
        for (int r = 0; r < UP_MAX; ++r)
            if (r < u && t0 + u * tx + r < p.Lout) yr[r] = acc[r][c] + bv;
    }
}

int n_groups(int cout) { return (cout + CO - 1) / CO; }

template <int KT, bool DIL1>
int launch_conv_t(const ConvParams& p, dim3 grid, size_t smem, cudaStream_t s) {
    static PerDeviceSlot<bool> attr_once;
    bool& attr_set = attr_once.get();
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(hg_conv_kernel<KT, DIL1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set = true;
    }
    hg_conv_kernel<KT, DIL1><<<grid, HG_THREADS, smem, s>>>(p);
    ES_LAUNCH_OK();
    return 0;
}

int launch_conv(const ConvParams& p, int B, cudaStream_t s) {
    const int n_cg = n_groups(p.Cout);
    ES_CHECK(n_cg >= 1 && n_cg <= 16 && (n_cg & (n_cg - 1)) == 0, "output channels must split into 1, 2, 4, 8 or 16 groups of 8");
    ES_CHECK(p.Cin % CI_CHUNK == 0, "input channels must be a multiple of 8");
    ES_CHECK(p.K >= 1 && p.K <= HG_MAX_K && (p.K & 1), "odd kernel size up to 15");
    const int TX = HG_THREADS / n_cg, TL = TX * PT;
    const int in_ld = (TL + (p.K - 1) * p.dil + 3) & ~3;
    const size_t smem = ((size_t)CI_CHUNK * in_ld + (size_t)CI_CHUNK * p.K * n_cg * CO) * sizeof(float);
    ES_CHECK(smem <= 160 * 1024, "tile does not fit shared memory");
    dim3 grid((unsigned)((p.Lout + TL - 1) / TL), (unsigned)B);
    const bool d1 = p.dil == 1;
    switch (p.K) {
        case 3: return d1 ? launch_conv_t<3, true>(p, grid, smem, s) : launch_conv_t<3, false>(p, grid, smem, s);
        case 7: return d1 ? launch_conv_t<7, true>(p, grid, smem, s) : launch_conv_t<7, false>(p, grid, smem, s);
        case 11: return d1 ? launch_conv_t<11, true>(p, grid, smem, s) : launch_conv_t<11, false>(p, grid, smem, s);
        default: return launch_conv_t<0, false>(p, grid, smem, s);
    }
}

int launch_upsample(const ConvParams& p, int B, cudaStream_t s) {
    const int n_cg = n_groups(p.Cout);
    ES_CHECK(n_cg >= 1 && n_cg <= 16 && (n_cg & (n_cg - 1)) == 0, "output channels must split into 1, 2, 4, 8 or 16 groups of 8");
    ES_CHECK(p.Cin % CI_CHUNK == 0, "input channels must be a multiple of 8");
    ES_CHECK(p.K >= 1 && p.K <= HG_MAX_K && p.stride >= 1, "transposed kernel size up to 16");
    ES_CHECK(p.stride <= UP_MAX && p.pad >= 0 && p.pad < p.K, "upsampling rate up to 8");
    const int TX = HG_THREADS / n_cg, TP = TX * p.stride;
    const int in_ld = (TX + (p.stride - 1 + p.pad) / p.stride + (p.K - 1 - p.pad + p.stride - 1) / p.stride + 3) & ~3;
    const size_t smem = ((size_t)CI_CHUNK * in_ld + (size_t)CI_CHUNK * p.K * n_cg * CO) * sizeof(float);
    ES_CHECK(smem <= 160 * 1024, "tile does not fit shared memory");
    static PerDeviceSlot<bool> attr_once;
    bool& attr_set = attr_once.get();
    if (!attr_set) {
        ES_CUDA(cudaFuncSetAttribute(hg_upsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set = true;
    }
    dim3 grid((unsigned)((p.Lout + TP - 1) / TP), (unsigned)B);
    hg_upsample_kernel<<<grid, HG_THREADS, smem, s>>>(p);
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace
}  // namespace es

struct es_hifigan {
    es_hifigan_config_t cfg;
    es_hifigan_weights_t w;
    int total_up;          // product of the upsample rates
};

extern "C" {

int es_hifigan_create(const es_hifigan_config_t* cfg, const es_hifigan_weights_t* w, es_hifigan_t** out) {
    using namespace es;
    ES_CHECK(cfg && w && out, "null argument");
    ES_CHECK(cfg->n_mel >= 8 && cfg->n_mel % 8 == 0, "n_mel must be a multiple of 8");
    ES_CHECK(cfg->n_up >= 1 && cfg->n_up <= ES_HG_MAX_UPS && cfg->n_res >= 1 && cfg->n_res <= ES_HG_MAX_RES, "bad stage counts");
    ES_CHECK(cfg->initial_channel % (8 << cfg->n_up) == 0 && cfg->initial_channel <= 128, "initial channel count must stay a multiple of 8 through every halving");
    int up = 1;
    for (int i = 0; i < cfg->n_up; ++i) {
        ES_CHECK(cfg->up_rate[i] >= 1 && cfg->up_kernel[i] >= cfg->up_rate[i] && cfg->up_kernel[i] <= 16 &&
                 (cfg->up_kernel[i] - cfg->up_rate[i]) % 2 == 0, "bad upsample rate / kernel");
        up *= cfg->up_rate[i];
    }
    for (int j = 0; j < cfg->n_res; ++j) {
        ES_CHECK(cfg->res_kernel[j] >= 1 && cfg->res_kernel[j] <= 15 && (cfg->res_kernel[j] & 1), "resblock kernels must be odd, <= 15");
        for (int d = 0; d < 3; ++d) ES_CHECK(cfg->res_dilation[j][d] >= 1 && cfg->res_dilation[j][d] <= 16, "bad dilation");
    }
    es_hifigan* h = new (std::nothrow) es_hifigan;
    ES_CHECK(h, "out of memory");
    h->cfg = *cfg;
    h->w = *w;
    h->total_up = up;
    *out = h;
    return 0;
}

void es_hifigan_destroy(es_hifigan_t* h) { delete h; }

size_t es_hifigan_workspace_bytes(const es_hifigan_t* h, int B, int T) {
    if (!h || B <= 0 || T <= 0) return 0;
    // four rotating [B, C_i, L_i] buffers sized for the largest stage (C_i L_i is constant from the second stage on
    // for rate-2 stages; take the maximum)
    size_t mx = (size_t)h->cfg.initial_channel * T;
    long long L = T;
    int C = h->cfg.initial_channel;
    for (int i = 0; i < h->cfg.n_up; ++i) {
        L *= h->cfg.up_rate[i];
        C /= 2;
        const size_t v = (size_t)C * (size_t)L;
        if (v > mx) mx = v;
    }
    return 4 * ((mx * (size_t)B * sizeof(float) + 255) / 256 * 256) + 256;
}

int es_hifigan_forward(es_hifigan_t* h, void* stream, int B, int T, const float* mel, long long mel_sb, long long mel_sc,
                       long long mel_st, float* wav, void* workspace, size_t workspace_bytes) {
    using namespace es;
    ES_CHECK(h, "null vocoder");
    ES_CHECK(B >= 1 && B <= 65535 && T >= 1, "need 1 <= B <= 65535 and T >= 1 mel frames");
    ES_CHECK(mel && wav, "null tensor");
    ES_CHECK(workspace && workspace_bytes >= es_hifigan_workspace_bytes(h, B, T), "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const es_hifigan_config_t& c = h->cfg;
    size_t mx = (es_hifigan_workspace_bytes(h, B, T) - 256) / 4;
    char* base = reinterpret_cast<char*>((reinterpret_cast<size_t>(workspace) + 255) / 256 * 256);
    float* buf[4];
    for (int i = 0; i < 4; ++i) buf[i] = reinterpret_cast<float*>(base + i * mx);

    auto conv = [&](const float* x, long long sb, long long sc, long long st, const es_hg_conv_w_t& w, int Cin, int Cout,
                    int Cout_store, int L, int K, int dil, float slope, const float* res, float scale, int accumulate,
                    int tanh_out, float* y) {
        ConvParams p;
        p.x = x; p.sb = sb; p.sc = sc; p.st = st; p.w = w.w; p.bias = w.b; p.res = res; p.y = y;
        p.Cin = Cin; p.Cout = Cout; p.Cout_store = Cout_store; p.Lin = L; p.Lout = L; p.K = K; p.dil = dil;
        p.pad = (K * dil - dil) / 2;                       // get_padding (hifigan/models.py:14-15)
        p.stride = 1; p.slope_in = slope; p.out_scale = scale; p.accumulate = accumulate; p.tanh_out = tanh_out;
        ES_CHECK(w.w && w.b, "missing weights");
        return launch_conv(p, B, s);
    };

    // conv_pre                                                                          hifigan/models.py:90, :112
    int C = c.initial_channel;
    long long L = T;
    if (conv(mel, mel_sb, mel_sc, mel_st, h->w.conv_pre, c.n_mel, C, C, (int)L, 7, 1, 1.f, nullptr, 1.f, 0, 0, buf[0])) return 1;
    float* x = buf[0];
    for (int i = 0; i < c.n_up; ++i) {
        // x = ups[i](leaky_relu(x, 0.1))                                                :114-115
        const int Cn = C / 2, u = c.up_rate[i], K = c.up_kernel[i];
        const long long Ln = L * u;
        ES_CHECK(Ln < 0x7fffffffLL, "waveform too long");
        float* up = (x == buf[0]) ? buf[1] : buf[0];
        {
            ConvParams p;
            p.x = x; p.sb = (long long)C * L; p.sc = L; p.st = 1; p.w = h->w.ups[i].w; p.bias = h->w.ups[i].b; p.res = nullptr;
            p.y = up; p.Cin = C; p.Cout = Cn; p.Cout_store = Cn; p.Lin = (int)L; p.Lout = (int)Ln; p.K = K; p.dil = 1;
            p.pad = (K - u) / 2; p.stride = u; p.slope_in = 0.1f; p.out_scale = 1.f; p.accumulate = 0; p.tanh_out = 0;
            ES_CHECK(p.w && p.bias, "missing upsample weights");
            if (launch_upsample(p, B, s)) return 1;
        }
        C = Cn; L = Ln;
        // xs = mean_j resblock_j(up)                                                    :116-122
        float* xs = (up == buf[0]) ? buf[1] : buf[0];         // the stage input x is dead now
        float* r = buf[2];
        float* t = buf[3];
        const long long sb = (long long)C * L;
        for (int j = 0; j < c.n_res; ++j) {
            const int K = c.res_kernel[j];
            const float* cur = up;                            // ResBlock1.forward (:45-52): x runs through 3 pairs
            for (int d = 0; d < 3; ++d) {
                const es_hg_conv_w_t& w1 = h->w.res[i][j].convs1[d];
                const es_hg_conv_w_t& w2 = h->w.res[i][j].convs2[d];
                // xt = c1(leaky_relu(x, 0.1))  (dilated)
                if (conv(cur, sb, L, 1, w1, C, C, C, (int)L, K, c.res_dilation[j][d], 0.1f, nullptr, 1.f, 0, 0, t)) return 1;
                // x = c2(leaky_relu(xt, 0.1)) + x; the last pair writes (x / n_res) into the stage sum instead
                const bool last = d == 2;
                if (conv(t, sb, L, 1, w2, C, C, C, (int)L, K, 1, 0.1f, cur, last ? 1.f / (float)c.n_res : 1.f,
                         last && j > 0 ? 1 : 0, 0, last ? xs : r)) return 1;
                cur = r;
            }
        }
        x = xs;
    }
    // wav = tanh(conv_post(leaky_relu(x)))   (default slope 0.01)                       :123-125
    es_hg_conv_w_t post = h->w.conv_post;
    return conv(x, (long long)C * L, L, 1, post, C, 1, 1, (int)L, 7, 1, 0.01f, nullptr, 1.f, 0, 1, wav);
}

}  // extern "C"

// Training-step pieces that sit AROUND the network (BASELINE configs[4]):
//
//   es_loss        EfficientSpeech.loss + the weighted total of training_step (model.py:167-217): masked L1 on the mel,
//                  masked MSE on pitch / energy, masked MSE on log(duration + 1); weights 10 / 2 / 2 / 1.  One pass also
//                  writes the gradients of the total with respect to the four predictions -- the seeds of the backward.
//   es_adamw_step  torch.optim.AdamW (model.py:279-283: lr 1e-3, weight_decay 1e-6, betas / eps at their defaults) as ONE
//                  kernel over the flat parameter / gradient / moment buffers.
//
// The backward of the network itself is NOT built: these two kernels plus one NCCL all-reduce of the
// flat gradient (sharding.allreduce_flat) are the pieces of the step that do not depend on it.
//
// Reductions are deterministic: per-block partial sums in a fixed order, then one block adds the partials in index
// order.  fp32 accumulation inside a block, double across blocks.
#include "es_common.cuh"

namespace es {
namespace {

constexpr int LT = 256;
constexpr int LOSS_BLOCKS = 592;          // 4 per SM; partials [LOSS_BLOCKS][4] doubles

__device__ __forceinline__ float block_sum(float v, float* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < LT / 32) t = sh[threadIdx.x];
    if (threadIdx.x < 32) t = warp_sum(t);
    return t;                              // valid in thread 0
}

// counts[0] = number of valid mel elements, counts[1] = number of valid phonemes
__global__ void __launch_bounds__(LT)
loss_count_kernel(const int32_t* __restrict__ mel_len, const uint8_t* __restrict__ pmask, int B, int N, int T, int n_mel,
                  double* __restrict__ counts) {
    __shared__ float sh[LT / 32];
    float cm = 0.f, cp = 0.f;
    for (int b = threadIdx.x; b < B; b += LT) cm += (float)min(max(__ldg(mel_len + b), 0), T);
    for (int i = threadIdx.x; i < B * N; i += LT) cp += pmask ? (pmask[i] ? 0.f : 1.f) : 1.f;
    const float a = block_sum(cm, sh);
    const float c = block_sum(cp, sh);
    if (threadIdx.x == 0) { counts[0] = (double)a * n_mel; counts[1] = (double)c; }
}

struct LossParams {
    int B, N, T, n_mel;
    const float* mel_pred; const float* mel_tgt; const int32_t* mel_len;
    const float* pitch_pred; const float* energy_pred; const float* dur_pred;
    const float* pitch; const float* energy; const int32_t* duration; const uint8_t* pmask;
    float* d_mel; float* d_pitch; float* d_energy; float* d_dur;
    const double* counts; double* partials;
};

__global__ void __launch_bounds__(LT)
loss_main_kernel(const LossParams p) {
    __shared__ float sh[LT / 32];
    const float inv_mel = p.counts[0] > 0 ? (float)(1.0 / p.counts[0]) : 0.f;
    const float inv_ph = p.counts[1] > 0 ? (float)(1.0 / p.counts[1]) : 0.f;
    // ---- mel: L1 over frames t < mel_len[b]                                     model.py:182-186
    float s_mel = 0.f;
    const long long n_rows = (long long)p.B * p.T;
    const int q = p.n_mel / 4;
    for (long long i = (long long)blockIdx.x * LT + threadIdx.x; i < n_rows * q; i += (long long)gridDim.x * LT) {
        const long long row = i / q;
        const int b = (int)(row / p.T), t = (int)(row - (long long)b * p.T);
        const bool valid = t < __ldg(p.mel_len + b);
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p.mel_pred) + i);
            const float4 y = __ldg(reinterpret_cast<const float4*>(p.mel_tgt) + i);
            const float d0 = a.x - y.x, d1 = a.y - y.y, d2 = a.z - y.z, d3 = a.w - y.w;
            s_mel += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
            const float w = 10.f * inv_mel;                                       // d(10 * mean |d|) / d pred
            g = make_float4(d0 > 0 ? w : (d0 < 0 ? -w : 0.f), d1 > 0 ? w : (d1 < 0 ? -w : 0.f),
                            d2 > 0 ? w : (d2 < 0 ? -w : 0.f), d3 > 0 ? w : (d3 < 0 ? -w : 0.f));
        }
        if (p.d_mel) reinterpret_cast<float4*>(p.d_mel)[i] = g;
    }
    // ---- pitch / energy MSE, duration MSE in the log domain, over un-padded phonemes   model.py:188-207
    float s_p = 0.f, s_e = 0.f, s_d = 0.f;
    for (int i = blockIdx.x * LT + threadIdx.x; i < p.B * p.N; i += gridDim.x * LT) {
        const bool valid = p.pmask ? p.pmask[i] == 0 : true;
        float gp = 0.f, ge = 0.f, gd = 0.f;
        if (valid) {
            const float dp = __ldg(p.pitch_pred + i) - __ldg(p.pitch + i);
            const float de = __ldg(p.energy_pred + i) - __ldg(p.energy + i);
            const float dpred = __ldg(p.dur_pred + i);
            const float dd = logf(dpred + 1.f) - logf((float)__ldg(p.duration + i) + 1.f);
            s_p = fmaf(dp, dp, s_p); s_e = fmaf(de, de, s_e); s_d = fmaf(dd, dd, s_d);
            gp = 2.f * 2.f * dp * inv_ph;                                         // weight 2, d mean(d^2) = 2 d / n
            ge = 2.f * 2.f * de * inv_ph;
            gd = 2.f * dd * inv_ph / (dpred + 1.f);                               // weight 1, chain through log(pred + 1)
        }
        if (p.d_pitch) p.d_pitch[i] = gp;
        if (p.d_energy) p.d_energy[i] = ge;
        if (p.d_dur) p.d_dur[i] = gd;
    }
    const float a = block_sum(s_mel, sh), b2 = block_sum(s_p, sh), c = block_sum(s_e, sh), d = block_sum(s_d, sh);
    if (threadIdx.x == 0) {
        double* o = p.partials + 4 * blockIdx.x;
        o[0] = a; o[1] = b2; o[2] = c; o[3] = d;
    }
}

// losses[0..4] = total, mel, pitch, energy, duration                               model.py:215-217
__global__ void loss_final_kernel(const double* __restrict__ partials, const double* __restrict__ counts, int n_blocks,
                                  float* __restrict__ losses) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s[4] = {0, 0, 0, 0};
    for (int k = 0; k < n_blocks; ++k)
        for (int j = 0; j < 4; ++j) s[j] += partials[4 * k + j];
    const double mel = counts[0] > 0 ? s[0] / counts[0] : 0.0;
    const double inv = counts[1] > 0 ? 1.0 / counts[1] : 0.0;
    const double pitch = s[1] * inv, energy = s[2] * inv, dur = s[3] * inv;
    losses[0] = (float)(10.0 * mel + 2.0 * pitch + 2.0 * energy + dur);
    losses[1] = (float)mel; losses[2] = (float)pitch; losses[3] = (float)energy; losses[4] = (float)dur;
}

// torch.optim.AdamW, single-tensor semantics (decoupled weight decay first, lerp moment update, bias corrections passed
// from the host where torch computes them in double)
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, size_t n,
             float lr, float beta1, float beta2, float eps, float weight_decay, float step_size, float bc2_sqrt) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float gi = g[i];
        float pi = p[i];
        pi = pi * (1.f - lr * weight_decay);                                      // param.mul_(1 - lr * weight_decay)
        const float mi = m[i] + (1.f - beta1) * (gi - m[i]);                      // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(1.f - beta2, gi * gi, v[i] * beta2);                // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi = pi + (-step_size * mi) / denom;                                      // param.addcdiv_(exp_avg, denom, value=-step_size)
        p[i] = pi; m[i] = mi; v[i] = vi;
    }
}

}  // namespace
}  // namespace es

extern "C" {

size_t es_loss_workspace_bytes(void) { return (size_t)(es::LOSS_BLOCKS * 4 + 2) * sizeof(double) + 256; }

int es_loss(void* stream, int B, int N, int T, int n_mel, const float* mel_pred, const float* mel_tgt, const int32_t* mel_len,
            const float* pitch_pred, const float* energy_pred, const float* dur_pred, const float* pitch, const float* energy,
            const int32_t* duration, const uint8_t* phoneme_mask, float* losses, float* d_mel, float* d_pitch, float* d_energy,
            float* d_dur, void* workspace, size_t workspace_bytes) {
    using namespace es;
    ES_CHECK(B >= 1 && N >= 1 && T >= 1 && n_mel >= 4 && n_mel % 4 == 0, "bad shape (n_mel must be a multiple of 4)");
    ES_CHECK(mel_pred && mel_tgt && mel_len && pitch_pred && energy_pred && dur_pred && pitch && energy && duration && losses, "null tensor");
    ES_CHECK(workspace && workspace_bytes >= es_loss_workspace_bytes(), "workspace too small");
    ES_CHECK((reinterpret_cast<size_t>(mel_pred) & 15) == 0 && (reinterpret_cast<size_t>(mel_tgt) & 15) == 0 &&
             (!d_mel || (reinterpret_cast<size_t>(d_mel) & 15) == 0), "mel buffers must be 16-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    double* counts = reinterpret_cast<double*>((reinterpret_cast<size_t>(workspace) + 255) / 256 * 256);
    double* partials = counts + 2;
    loss_count_kernel<<<1, LT, 0, s>>>(mel_len, phoneme_mask, B, N, T, n_mel, counts);
    ES_LAUNCH_OK();
    LossParams p;
    p.B = B; p.N = N; p.T = T; p.n_mel = n_mel; p.mel_pred = mel_pred; p.mel_tgt = mel_tgt; p.mel_len = mel_len;
    p.pitch_pred = pitch_pred; p.energy_pred = energy_pred; p.dur_pred = dur_pred; p.pitch = pitch; p.energy = energy;
    p.duration = duration; p.pmask = phoneme_mask; p.d_mel = d_mel; p.d_pitch = d_pitch; p.d_energy = d_energy; p.d_dur = d_dur;
    p.counts = counts; p.partials = partials;
    loss_main_kernel<<<LOSS_BLOCKS, LT, 0, s>>>(p);
    ES_LAUNCH_OK();
    loss_final_kernel<<<1, 32, 0, s>>>(partials, counts, LOSS_BLOCKS, losses);
    ES_LAUNCH_OK();
    return 0;
}

int es_adamw_step(void* stream, size_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr,
                  float beta1, float beta2, float eps, float weight_decay, float step_size, float bias_correction2_sqrt) {
    using namespace es;
    ES_CHECK(param && grad && exp_avg && exp_avg_sq, "null tensor");
    ES_CHECK(bias_correction2_sqrt > 0.f, "bias correction must be positive (step >= 1)");
    if (n == 0) return 0;
    size_t blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    adamw_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                                  eps, weight_decay, step_size, bias_correction2_sqrt);
    ES_LAUNCH_OK();
    return 0;
}

}  // extern "C"

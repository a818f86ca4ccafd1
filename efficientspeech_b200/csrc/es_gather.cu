// Length regulator as an index map (FeatureUpsampler.forward, layers/networks.py:228-258).
//
// The decoder's input projection skip = LN(tanh(Linear(features))) (networks.py:292) is row-wise, and
// features[b, t] = fused4[b, src(b, t)] is a row gather, so the two commute:
//     skip[b, t] = P[b * N + src(b, t)],   P = LN(tanh(Linear(fused4)))  computed once per PHONEME
// (B*N rows instead of B*T frames: 6x fewer rows at the headline shape).  Frames past mel_len[b] are
// zero rows of `features` in the reference (pad_sequence, networks.py:247-252), i.e. the constant
// row LN(tanh(bias)); it is stored as one extra row P[B*N].
//
//   frame_source_kernel  src[b*T + t] = b*N + upper_bound(cum[b,:], t)   (t <  mel_len[b])
//                                     = B*N                              (t >= mel_len[b])
//                        integer work, bit-exact with torch.repeat_interleave; CTA (0,0) also writes
//                        the pad row.  Consumed by the decoder kernel's gathered loads
//                        (es_umma_dec.cu) or by gather_rows_kernel.
//   gather_rows_kernel   Y[r, :] = P[src[r], :], 128-bit accesses (decoders whose layer kernel has no
//                        gathered-load variant: dx2 = 256, fp32 SIMT mode).
#include <cuda_fp16.h>

#include "es_common.cuh"
#include "es_kernels.cuh"

namespace es {
namespace {

constexpr int FS_THREADS = 256;
constexpr int FS_MAX_SMEM_N = 8192;      // prefix sums staged in shared memory up to this many phonemes

// ---- ragged scheduling -------------------------------------------------------------------------------------------------
// A decoder layer is a 5-tap convolution along time: through L layers a frame depends on frames within 2 L of it.  The
// reference runs the decoder over every padded frame and then zeroes the frames past mel_len (networks.py:424-427), so
// tiles that start at or beyond mel_len[b] + halo (halo = 2 L) cannot reach a frame anyone reads.  frame_source_kernel
// therefore also (a) lists the other tiles -- utterance by utterance, so consecutive CTAs of the decoder kernels still
// stream consecutive memory -- for those kernels to walk instead of the dense B x ceil(T / TM) grid (CTA (0,0): one
// block-wide scan per 256 utterances), and (b) zero-fills the mel frames no listed tile covers.  No extra launch.
struct RaggedPlan {
    int2* tiles;          // out: (b, t0) per scheduled tile, or null (ragged scheduling off)
    int* count;           // out: number of scheduled tiles
    int TM, halo;         // tile size in frames; reach of the decoder in frames (2 per layer)
    float* mel;           // frames past the scheduled tiles are zero-filled here ...
    int n_mel;            // ... n_mel (% 4 == 0) floats per frame
};
__device__ __forceinline__ int tiles_of(int valid, int T, int TM, int halo) {
    const int ext = min(T, valid + halo);
    return valid > 0 ? (ext + TM - 1) / TM : 0;                 // an utterance without frames needs no tile at all
}

__global__ void __launch_bounds__(FS_THREADS)
frame_source_kernel(const int32_t* __restrict__ cum, const int32_t* __restrict__ valid_len,
                    int32_t* __restrict__ src, int B, int N, int T,
                    const float* __restrict__ bias, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                    int C, float* __restrict__ pad_row, const RaggedPlan rp) {
    extern __shared__ int scum[];
    const int b = blockIdx.y, t = blockIdx.x * FS_THREADS + threadIdx.x;
    const int32_t* c = cum + (size_t)b * N;
    const bool in_smem = N <= FS_MAX_SMEM_N;
    if (in_smem) {
        for (int k = threadIdx.x; k < N; k += FS_THREADS) scum[k] = __ldg(c + k);
        __syncthreads();
    }
    if (t < T) {
        int s = B * N;                                   // pad row
        if (t < __ldg(valid_len + b)) {
            int lo = 0, hi = N;                          // first index with cum > t
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const int cv = in_smem ? scum[mid] : __ldg(c + mid);
                if (cv > t) hi = mid; else lo = mid + 1;
            }
            if (lo < N) s = b * N + lo;
        }
        src[(size_t)b * T + t] = s;
        if (rp.tiles && t >= tiles_of(__ldg(valid_len + b), T, rp.TM, rp.halo) * rp.TM) {
            float4* y = reinterpret_cast<float4*>(rp.mel + ((size_t)b * T + t) * rp.n_mel);
            for (int k = 0; k < rp.n_mel / 4; ++k) y[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    // pad row = LN(tanh(0 * W + bias)): one warp, two-pass statistics
    if (pad_row && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 32) {
        const int lane = threadIdx.x;
        float v[8];                                      // C <= 256
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ch = lane + 32 * k;
            v[k] = ch < C ? tanhf(__ldg(bias + ch)) : 0.f;
            sum += v[k];
        }
        const float mean = warp_sum(sum) / (float)C;
        float var = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ch = lane + 32 * k;
            const float dlt = ch < C ? v[k] - mean : 0.f;
            var = fmaf(dlt, dlt, var);
        }
        const float rstd = rsqrtf(warp_sum(var) / (float)C + kLnEps);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ch = lane + 32 * k;
            if (ch < C) pad_row[ch] = (v[k] - mean) * rstd * __ldg(ln_g + ch) + __ldg(ln_b + ch);
        }
    }
    // the tile list: exclusive scan of the per-utterance tile counts, 256 utterances per round
    if (rp.tiles && blockIdx.x == 0 && blockIdx.y == 0) {
        __shared__ int warp_sums[FS_THREADS / 32];
        __shared__ int base;
        if (threadIdx.x == 0) base = 0;
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int b0 = 0; b0 < B; b0 += FS_THREADS) {
            const int u = b0 + threadIdx.x;
            const int n = u < B ? tiles_of(__ldg(valid_len + u), T, rp.TM, rp.halo) : 0;
            int incl = n;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) warp_sums[warp] = incl;
            __syncthreads();
            int before = base;
            for (int w = 0; w < warp; ++w) before += warp_sums[w];
            const int start = before + incl - n;
            for (int k = 0; k < n; ++k) rp.tiles[start + k] = make_int2(u, k * rp.TM);
            __syncthreads();
            if (threadIdx.x == FS_THREADS - 1) base = before + incl;
            __syncthreads();
        }
        if (threadIdx.x == 0) *rp.count = base;
    }
}

// 8 rows per 256-thread CTA step, grid-stride; every lane moves 16 bytes per step
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float* __restrict__ P, const int32_t* __restrict__ src, float* __restrict__ Y,
                   long long rows, int C4) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (long long r = (long long)blockIdx.x * 8 + w; r < rows; r += (long long)gridDim.x * 8) {
        const float4* sp = reinterpret_cast<const float4*>(P) + (size_t)__ldg(src + r) * C4;
        float4* dp = reinterpret_cast<float4*>(Y) + (size_t)r * C4;
        for (int i = lane; i < C4; i += 32) dp[i] = __ldg(sp + i);
    }
}

}  // namespace

int launch_frame_source(const int32_t* cum, const int32_t* valid_len, int32_t* src, int B, int N, int T,
                        const float* bias, const float* ln_g, const float* ln_b, int C, float* pad_row,
                        cudaStream_t s, int2* tiles, int* tile_count, int tile_frames, int halo, float* mel, int n_mel) {
    ES_CHECK(cum && valid_len && src, "null tensor");
    ES_CHECK(B >= 1 && B <= 65535 && N >= 1 && T >= 1, "bad shape");
    ES_CHECK((long long)B * N < 0x7fffffffLL, "B * N overflows the row index");
    ES_CHECK(!pad_row || (bias && ln_g && ln_b && C >= 32 && C <= 256), "pad row needs bias / LayerNorm over <= 256 channels");
    ES_CHECK(!tiles || (tile_count && mel && tile_frames >= 1 && halo >= 0 && n_mel % 4 == 0), "bad ragged plan");
    const dim3 grid((unsigned)((T + FS_THREADS - 1) / FS_THREADS), (unsigned)B);
    const size_t smem = N <= FS_MAX_SMEM_N ? (size_t)N * sizeof(int) : 0;
    RaggedPlan rp;
    rp.tiles = tiles; rp.count = tile_count; rp.TM = tile_frames; rp.halo = halo; rp.mel = mel; rp.n_mel = n_mel;
    frame_source_kernel<<<grid, FS_THREADS, smem, s>>>(cum, valid_len, src, B, N, T, bias, ln_g, ln_b, C, pad_row, rp);
    ES_LAUNCH_OK();
    return 0;
}

int launch_gather_rows(const float* P, const int32_t* src, float* Y, long long rows, int C, cudaStream_t s) {
    ES_CHECK(P && src && Y && C % 4 == 0, "bad argument");
    if (rows <= 0) return 0;
    long long blocks = (rows + 7) / 8;
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    gather_rows_kernel<<<(unsigned)blocks, 256, 0, s>>>(P, src, Y, rows, C / 4);
    ES_LAUNCH_OK();
    return 0;
}

// fp32 -> fp16 (round to nearest even), 8 elements per thread: the opt-in half-precision copy of the mel that halves the
// device -> host bytes of a PCIe-bound consumer (bench.py e2e); the fp32 mel stays the API's output.
namespace {
__global__ void __launch_bounds__(256)
cast_f32_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
    const size_t i = ((size_t)blockIdx.x * 256 + threadIdx.x) * 8;
    if (i + 8 <= n) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
        const float4 b = __ldg(reinterpret_cast<const float4*>(src + i) + 1);
        __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w), __floats2half2_rn(b.x, b.y), __floats2half2_rn(b.z, b.w)};
        *reinterpret_cast<uint4*>(dst + i) = *reinterpret_cast<const uint4*>(h);
    } else {
        for (size_t k = i; k < n; ++k) dst[k] = __float2half_rn(src[k]);
    }
}
}  // namespace

int launch_cast_f32_f16(const float* src, void* dst, size_t n, cudaStream_t s) {
    ES_CHECK(src && dst, "null tensor");
    ES_CHECK((reinterpret_cast<size_t>(src) & 15) == 0 && (reinterpret_cast<size_t>(dst) & 15) == 0, "buffers must be 16-byte aligned");
    if (n == 0) return 0;
    const size_t blocks = (n + 2047) / 2048;
    cast_f32_f16_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, static_cast<__half*>(dst), n);
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace es

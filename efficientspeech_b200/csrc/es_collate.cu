// Batch collation on the device: LJSpeechDataModule.collate_fn (datamodule.py:29-76) + get_mask_from_lengths
// (utils/tools.py:43-51) for the acoustic model's inputs.  The step BEFORE the hot path.
//
// The reference sorts the utterances of a batch by decreasing phoneme count on the host (np.argsort(-len),
// datamodule.py:31-32), pads every per-phoneme array with zeros to the longest utterance (pad_1D, utils/tools.py:262-272)
// and builds the boolean padding mask ids >= len.  Here the ragged arrays arrive concatenated (CSR offsets) and two
// kernels do the same:
//
//   collate_rank_kernel   perm = stable argsort of -len: rank(i) = #{j: len_j > len_i} + #{j < i: len_j == len_i}
//                         (integer compares over lengths staged in shared memory, one thread per utterance; ties keep
//                         their input order -- the reference's introsort leaves tie order unspecified, any order is a
//                         permutation of equal-length rows)
//   collate_fill_kernel   row r <- utterance perm[r]: coalesced copies + zero padding of phoneme / pitch / energy /
//                         duration, mask, phoneme_len, and mel_len[r] = sum of durations (what the dataset stores as
//                         the mel length, preprocessor output; warp-shuffle reduction)
//
// Pure integer / copy work: results are bit-exact by construction and tested against the oracle's restatement.
#include "es_common.cuh"
#include "es_kernels.cuh"

namespace es {
namespace {

constexpr int CR_THREADS = 256;

__global__ void __launch_bounds__(CR_THREADS)
collate_rank_kernel(const int32_t* __restrict__ offsets, int32_t* __restrict__ perm, int B) {
    extern __shared__ int slen[];
    for (int k = threadIdx.x; k < B; k += CR_THREADS) slen[k] = __ldg(offsets + k + 1) - __ldg(offsets + k);
    __syncthreads();
    for (int i = blockIdx.x * CR_THREADS + threadIdx.x; i < B; i += gridDim.x * CR_THREADS) {
        const int li = slen[i];
        int rank = 0;
        for (int j = 0; j < B; ++j) {
            const int lj = slen[j];
            rank += (lj > li) || (lj == li && j < i);
        }
        perm[rank] = i;
    }
}

__global__ void __launch_bounds__(128)
collate_fill_kernel(const int32_t* __restrict__ offsets, const int32_t* __restrict__ perm,
                    const int32_t* __restrict__ ph_flat, const float* __restrict__ pitch_flat,
                    const float* __restrict__ energy_flat, const int32_t* __restrict__ dur_flat,
                    int32_t* __restrict__ phoneme, uint8_t* __restrict__ mask, int32_t* __restrict__ phoneme_len,
                    float* __restrict__ pitch, float* __restrict__ energy, int32_t* __restrict__ duration,
                    int32_t* __restrict__ mel_len, int N) {
    const int r = blockIdx.x;
    const int u = __ldg(perm + r);
    const int beg = __ldg(offsets + u), len = __ldg(offsets + u + 1) - beg;
    int dsum = 0;
    for (int n = threadIdx.x; n < N; n += 128) {
        const bool in = n < len;
        const size_t o = (size_t)r * N + n;
        phoneme[o] = in ? __ldg(ph_flat + beg + n) : 0;
        mask[o] = in ? 0 : 1;                                  // True = padding (ids >= len, utils/tools.py:49)
        if (pitch) pitch[o] = in ? __ldg(pitch_flat + beg + n) : 0.f;
        if (energy) energy[o] = in ? __ldg(energy_flat + beg + n) : 0.f;
        if (duration) {
            const int d = in ? __ldg(dur_flat + beg + n) : 0;
            duration[o] = d;
            dsum += d;
        }
    }
    if (threadIdx.x == 0) phoneme_len[r] = len;
    if (mel_len) {
        __shared__ int part[4];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = dsum;
        __syncthreads();
        if (threadIdx.x == 0) mel_len[r] = part[0] + part[1] + part[2] + part[3];
    }
}

}  // namespace

int launch_collate(int B, int N, const int32_t* offsets, const int32_t* ph_flat, const float* pitch_flat,
                   const float* energy_flat, const int32_t* dur_flat, int32_t* perm, int32_t* phoneme, uint8_t* mask,
                   int32_t* phoneme_len, float* pitch, float* energy, int32_t* duration, int32_t* mel_len, cudaStream_t s) {
    ES_CHECK(B >= 1 && B <= 12000 && N >= 1, "need 1 <= B <= 12000 utterances and N >= 1");
    ES_CHECK(offsets && ph_flat && perm && phoneme && mask && phoneme_len, "null tensor");
    ES_CHECK((!pitch || pitch_flat) && (!energy || energy_flat) && (!duration || dur_flat), "output without its input");
    ES_CHECK(!mel_len || duration, "mel_len is the sum of the durations");
    const int blocks = (B + CR_THREADS - 1) / CR_THREADS;
    collate_rank_kernel<<<blocks, CR_THREADS, (size_t)B * sizeof(int), s>>>(offsets, perm, B);
    ES_LAUNCH_OK();
    collate_fill_kernel<<<B, 128, 0, s>>>(offsets, perm, ph_flat, pitch_flat, energy_flat, dur_flat, phoneme, mask,
                                          phoneme_len, pitch, energy, duration, mel_len, N);
    ES_LAUNCH_OK();
    return 0;
}

}  // namespace es

// Launch wrappers of the non-GEMM kernels (es_encoder.cu) and of the tcgen05 path (es_umma.cu).
#pragma once
#include "es_common.cuh"

namespace es {

int launch_embed_merge(const int32_t* ids, const float* tab, float* out, int B, int N, int C, int k,
                       int n_symbols, cudaStream_t s);
int launch_pool_mask(const uint8_t* mask, uint8_t* out, int B, int N, int n1, int pool, cudaStream_t s);
int launch_attention(const float* qkv, float* out, int B, int n, int C, int H, float scale, cudaStream_t s);
int launch_fuse(const float* f0, const float* f1, const float* a0, const float* g, const float* gb,
                const float* cst, const uint8_t* mask, float* out, int B, int N, int n1, int d, int k,
                cudaStream_t s);
int launch_variance_scan(const float* fused, const float* dur_feat, const float* pitch_pred,
                         const float* energy_pred, const float* dur_pred, const float* pitch_tgt,
                         const float* energy_tgt, const int32_t* dur_tgt, const uint8_t* mask,
                         const es_predictor_w_t& pw, const es_predictor_w_t& ew, float* fused4,
                         int32_t* dur_int, int32_t* dur_cum, int32_t* mel_len, int B, int N, int d,
                         cudaStream_t s);
int launch_length_regulate(const float* fused4, const int32_t* cum, const uint8_t* pmask, float* feats,
                           uint8_t* fmask, int32_t* src, int B, int N, int T, int C, cudaStream_t s);

// length regulator as an index map + row gather (es_gather.cu)
int launch_frame_source(const int32_t* cum, const int32_t* valid_len, int32_t* src, int B, int N, int T,
                        const float* bias, const float* ln_g, const float* ln_b, int C, float* pad_row,
                        cudaStream_t s, int2* tiles = nullptr, int* tile_count = nullptr, int tile_frames = 0, int halo = 0,
                        float* mel = nullptr, int n_mel = 0);
// batch collation on the device (es_collate.cu)
int launch_collate(int B, int N, const int32_t* offsets, const int32_t* ph_flat, const float* pitch_flat,
                   const float* energy_flat, const int32_t* dur_flat, int32_t* perm, int32_t* phoneme, uint8_t* mask,
                   int32_t* phoneme_len, float* pitch, float* energy, int32_t* duration, int32_t* mel_len, cudaStream_t s);
int launch_cast_f32_f16(const float* src, void* dst, size_t n, cudaStream_t s);
// (ragged scheduling: with `tiles` given, launch_frame_source also lists the tiles of tile_frames frames that can reach a
// valid frame -- t0 < min(T, valid_len[b] + halo) -- and zero-fills the mel frames no listed tile covers)
int launch_gather_rows(const float* P, const int32_t* src, float* Y, long long rows, int C, cudaStream_t s);

// tcgen05 decoder kernel (es_umma_dec.cu)
bool umma_dec_supported(int C, int dw_k, int N);
int launch_umma_dec(int mode, int B, int T, int N, const float* X, const float* dw_w, const float* dw_b, const void* w_h16,
                    const float* bias, int act_tanh, const float* ln_g, const float* ln_b,
                    const float* res2, const float* ln2_g, const float* ln2_b, const int* zero_from,
                    float* Y, cudaStream_t s, const int2* tile_list = nullptr, const int* tile_count = nullptr);
// depthwise layer whose input rows (gather_x) and / or skip rows (gather_res2) are rows of a table addressed
// through the frame -> row map `src` [B*T] (launch_frame_source): X / res2 then point at the table, whose row
// `pad_id` (the largest index) serves the zero-padded frames
int launch_umma_dec_gathered(int B, int T, int N, const float* X, const float* dw_w, const float* dw_b,
                             const void* w_h16, const float* bias, int act_tanh, const float* ln_g, const float* ln_b,
                             const float* res2, const float* ln2_g, const float* ln2_b, const int* src, int pad_id,
                             int gather_x, int gather_res2, float* Y, cudaStream_t s, const int2* tile_list = nullptr,
                             const int* tile_count = nullptr);
int umma_dec_check_errors(cudaStream_t s);
// wide decoders (dx2 = 256): K-streamed tcgen05 kernel (es_umma_dec256.cu)
bool umma_dec256_supported(int K, int dw_k, int N, int mode);
int launch_umma_dec256(int mode, int B, int T, int K, int N, const float* X, const float* dw_w, const float* dw_b, const void* w_chunks,
                       const float* bias, int act_tanh, const float* ln_g, const float* ln_b,
                       const float* res2, const float* ln2_g, const float* ln2_b, const int* zero_from,
                       float* Y, cudaStream_t s, const int2* tile_list = nullptr, const int* tile_count = nullptr);
int launch_umma_dec256_gathered(int B, int T, const float* X, const float* dw_w, const float* dw_b, const void* w_chunks,
                                const float* bias, const float* ln_g, const float* ln_b,
                                const float* res2, const float* ln2_g, const float* ln2_b, const int* src, bool gx, bool gs,
                                float* Y, cudaStream_t s, const int2* tile_list = nullptr, const int* tile_count = nullptr);
void umma_dec_set_trace(long long* buf);
int* umma_err_flag();
// es_train_gemm.cu: tcgen05 split-16-bit GEMM of the training step; -1 = outside its envelope
int launch_train_gemm_tc(cudaStream_t s, int slices, int M, int N, int K, const float* A, int lda, long long sa, int ta,
                         const float* B, int ldb, long long sb, int tb, float* C, int ldc, long long sc, const float* bias, int k_chunk,
                         int wide_mask);

// tcgen05 row GEMM for the phoneme-side layers (es_umma_enc.cu); -1: outside its envelope
int launch_umma_rowgemm(const RowGemmParams& p, const void* w_h16, cudaStream_t s);
int launch_umma_rowgemm_batch(const RowGemmParams* ps, const void* const* w_h16, int count, cudaStream_t s);
// K-streamed row GEMM for the layers outside that envelope (es_umma_wide.cu); -1: not applicable
int dense_layout(int K, int Nout, int taps, int stride);   // 0 none, 1 resident taps image, 2/3 streamed units (NT 128/256)
int launch_umma_wide(const RowGemmParams& p, const void* w_units, cudaStream_t s, int nt_force = 0);
// the whole phoneme side (PhonemeEncoder.forward up to the upsampler) in one kernel for the tiny geometry
// (es_umma_phoneme.cu); -1: outside its envelope
bool umma_phoneme_supported(const es_config_t& cfg, const es_weights_t& w, int N);
void umma_phoneme_set_trace(long long* buf);   // debug: [2][128] (clock64, event code) pairs of CTA 0, or null
int launch_umma_phoneme(const es_config_t& cfg, const es_weights_t& w, int B, int N, int n1, int pool,
                        const int32_t* ids, const uint8_t* mask, const float* pitch_tgt, const float* energy_tgt,
                        const int32_t* dur_tgt, float* pitch_pred, float* energy_pred, float* dur_pred, float* fused4,
                        int32_t* dur_int, int32_t* dur_cum, int32_t* mel_len, float* sc_xm1, cudaStream_t s);
// tcgen05 attention (es_umma_attn.cu); -1: outside its envelope (n > 128 or C not in {32, 64})
int launch_umma_attention(const float* qkv, float* out, int B, int n, int C, int H, float scale, cudaStream_t s);

}  // namespace es

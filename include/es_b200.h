/*
 * es_b200.h -- C ABI of the B200-native EfficientSpeech acoustic forward path.
 *
 * The reference (roatienza/efficientspeech @ 218f62f) has no FFI of its own: the boundary
 * it exposes for this path is the Python nn.Module API
 *     layers/__init__.py:1        from .networks import PhonemeEncoder, MelDecoder, Phoneme2Mel
 *     model.py:132-147            construction of the three modules
 *     model.py:155-164            the forward that calls them
 * Each entry point below names the reference method it replaces.  The Python mirror of the
 * reference modules (efficientspeech_b200/modules.py) binds these with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - tensors are dense, row-major, channels-last: [B, n, C] fp32, ids/durations int32,
 *     masks uint8 (1 = padding, the reference's bool True);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call
 *     synchronises the device;
 *   - every function returns 0 on success, non-zero on error; es_last_error() returns a
 *     thread-local message (bad shape, unsupported geometry, CUDA launch failure ...).
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef ES_B200_H
#define ES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ES_ABI_VERSION 14
#define ES_MAX_ENC_BLOCKS 2
#define ES_MAX_DEC_LAYERS 24
#define ES_MAX_DEC_BLOCKS 8
#define ES_MAX_TAPS 9

/* Geometry; mirrors the ctor arguments of PhonemeEncoder / MelDecoder
 * (layers/networks.py:310-333, :264-270). */
typedef struct es_config {
    int32_t embed_dim;            /* 128 */
    int32_t dim;                  /* d = embed_dim / reduction */
    int32_t kernel_size;          /* encoder merge-conv kernel (3 | 5) */
    int32_t head;                 /* heads of block 0 (block 1 has 2x) */
    int32_t expansion;            /* MixFFN hidden multiplier */
    int32_t n_blocks;             /* decoder blocks */
    int32_t block_depth;          /* layers per decoder block */
    int32_t decoder_kernel_size;  /* depthwise kernel (5) */
    int32_t n_mel;                /* 80 */
    int32_t n_symbols;            /* 153 */
} es_config_t;

/* Folded, kernel-ready weights (fp32, device).  Built on the host by
 * efficientspeech_b200/packing.py from the reference state dict; layouts are
 * "[taps][K][Nout_padded]" (K-major rows, output channel contiguous, Nout padded to a
 * multiple of 32 with zeros) unless noted. */
typedef struct es_enc_block_w {
    const float* merge_w;    /* block 0: gather tables [k][n_symbols][C] = E . (W1x1 Wk)^T  (networks.py:54,65-66)
                                block 1: folded conv   [k'][Cin][C]      = W1x1 . Wk'        (networks.py:65-66) */
    const float* qkv_w;      /* [1][C][3HC]                 blocks.py:45 */
    const float* proj_w;     /* [1][HC][C]                  blocks.py:66 */
    const float* proj_b;     /* [C] */
    const float* ln1_g;      /* networks.py:73 */
    const float* ln1_b;
    const float* ffn1_w;     /* [3][C][hC] = conv3 . mlp1   blocks.py:23-25 */
    const float* ffn1_tapb;  /* [3][hC]    = Wc_tau . b1 (added only where tap tau is inside the sequence) */
    const float* ffn1_b;     /* [hC] conv bias */
    const float* ffn2_w;     /* [1][hC][C]                  blocks.py:28 */
    const float* ffn2_b;
    const float* ln2_g;      /* networks.py:80 */
    const float* ln2_b;
    /* split-fp16 images for the tcgen05 row GEMM, UMMA canonical K-major no-swizzle order
     * [taps][2 (hi, lo)][K/8][N][8] halves; NULL -> the fp32 SIMT kernel is used */
    const void*  merge_w_h16;   /* block 1 only */
    const void*  qkv_w_h16;
    const void*  proj_w_h16;
    const void*  ffn1_w_h16;
    const void*  ffn2_w_h16;
    /* block 1, merge conv with 3 taps and stride 2 (base: kernel_size 5 -> k' = 3): the same conv over PAIRED rows.
     * With X2[t] = [x[2t] | x[2t+1]] (a free view of [n][C] as [n/2][2C], n even),
     *   y[t] = W0 x[2t-1] + W1 x[2t] + W2 x[2t+1] = [0 | W0] X2[t-1] + [W1 | W2] X2[t]
     * is a stride-1 "same" conv with K = 2C, which the streamed-weight tcgen05 kernel runs (es_umma_wide.cu has no
     * strided form).  [3][2 Cin][C] (third tap zero) and its split-fp16 units; NULL -> strided fp32 SIMT kernel. */
    const float* merge2_w;
    const void*  merge2_w_h16;
} es_enc_block_w_t;

typedef struct es_predictor_w {     /* AcousticDecoder, networks.py:98-122,151-165 */
    const float* conv1_w;    /* [3][d][d] */
    const float* conv1_b;
    const float* ln1_g;
    const float* ln1_b;
    const float* conv2_w;    /* [3][d][d] */
    const float* conv2_b;
    const float* ln2_g;      /* used by the duration predictor only (networks.py:159) */
    const float* ln2_b;
    const float* lin_w;      /* [d] */
    const float* lin_b;      /* [1] */
    const float* bins;       /* [d-1]  (pitch / energy) or NULL */
    const float* table;      /* [d][d] (pitch / energy) or NULL */
    const void*  conv1_w_h16; /* [3][2][d/8][d][8] halves or NULL */
    const void*  conv2_w_h16;
} es_predictor_w_t;

typedef struct es_dec_layer_w {     /* networks.py:279-283 */
    const float* dw_w;       /* [k][dx2]   depthwise taps, channel contiguous */
    const float* dw_b;       /* [dx2] */
    const float* pw_w;       /* [1][dx2][dx2] */
    const float* pw_b;
    const float* ln_g;
    const float* ln_b;
    /* split-fp16 image of pw_w for the tcgen05 kernels in the UMMA canonical K-major no-swizzle
     * order.  dx2 == 128: [2 (hi, lo)][K/8][N][8] halves (resident weights, es_umma_dec.cu);
     * dx2 == 256: K-chunked [K/32][2 (hi, lo)][4][N][8] (streamed, es_umma_dec256.cu).  The same
     * convention holds for dproj_w_h16 / mel_w_h16.  NULL -> the SIMT kernel is used. */
    const void*  pw_w_h16;
} es_dec_layer_w_t;

typedef struct es_weights {
    es_enc_block_w_t enc[ES_MAX_ENC_BLOCKS];
    /* Fuse (networks.py:189-219), all linear maps folded: */
    const float* fuse_a0;    /* [d][d]        = Wf[:, :d] . Wm0            (K-major) */
    const float* fuse_g;     /* [k][2d][d]    = Wf[:, d:] . Wct_tau^T . Wm1 */
    const float* fuse_gb;    /* [k][d]        = Wf[:, d:] . Wct_tau^T . bm1 */
    const float* fuse_c;     /* [d]           = Wf[:, :d] bm0 + Wf[:, d:] bct + bf */
    /* tensor-core form of the same maps for d % 128 == 0 (base), both as streamed units with 128
     * output columns (es_dense_layout format 2): U = feat1 . [G_0 | .. | G_{k-1}] + [g_0 | .. | g_{k-1}]
     * for every half-rate position, then fused[t] = c + A0 feat0[t] + sum_{tau = t mod 2, ..} U_tau[(t - tau) / 2].
     * NULL -> the fp32 SIMT fuse kernel runs. */
    const void*  fuse_u_h16;   /* units image of [2d] -> [k*d] */
    const void*  fuse_a0_h16;  /* units image of [d] -> [d] */
    es_predictor_w_t pitch;
    es_predictor_w_t energy;
    es_predictor_w_t duration;
    /* MelDecoder (networks.py:272-288) */
    const float* dproj_w;    /* [1][dx4][dx2] */
    const float* dproj_b;
    const float* dproj_ln_g;
    const float* dproj_ln_b;
    es_dec_layer_w_t dec[ES_MAX_DEC_LAYERS];
    const float* blk_ln_g[ES_MAX_DEC_BLOCKS];
    const float* blk_ln_b[ES_MAX_DEC_BLOCKS];
    const float* mel_w;      /* [1][dx2][96] */
    const float* mel_b;      /* [96] */
    const void*  dproj_w_h16; /* canonical split-fp16 image of dproj_w ([2][dx4/8][dx2][8]) or NULL */
    const void*  mel_w_h16;   /* canonical split-fp16 image of mel_w   ([2][dx2/8][n_mel][8]) or NULL */
} es_weights_t;

typedef struct es_model es_model_t;   /* opaque */

int         es_abi_version(void);
const char* es_last_error(void);

/* Replaces PhonemeEncoder.__init__/MelDecoder.__init__/Phoneme2Mel.__init__ (networks.py:310-333,
 * :264-288, :407-413) as far as device state goes.  The weights struct is copied; the
 * buffers it points to stay owned by the caller and must outlive the model. */
int  es_model_create(const es_config_t* cfg, const es_weights_t* w, es_model_t** out);
void es_model_destroy(es_model_t* m);
/* 0: SIMT fp32 kernels everywhere; 1 (default): tcgen05 split-fp16 decoder layers where supported */
int  es_model_set_tensor_core(es_model_t* m, int enable);
/* 1 (default): es_encoder_forward runs the whole phoneme side as ONE kernel (one CTA per utterance, every
 * activation on chip) when the geometry allows it: dim 32, head 1, kernel_size 3, expansion 1 (tiny), 2 <= N <= 128.
 * 0: one launch per layer, as for every other geometry.  Same function, same parity bars. */
int  es_model_set_fused_phoneme(es_model_t* m, int enable);

/* 1 (default): es_decoder_forward_gathered with zero_padded_frames = 1 schedules only the decoder tiles that can reach a
 * valid frame (t0 < mel_len[b] + 2 L, L decoder layers) -- the ragged scheduling the reference's padding makes possible:
 * it computes every padded frame and then zeroes it (networks.py:424-427).  0: every tile, as the reference does.
 * Same results on every frame below mel_len, zeros above (tests compare the two bit for bit). */
int  es_model_set_ragged_schedule(es_model_t* m, int enable);

/* How es_decoder_forward_gathered joins the length regulator and the decoder (default ES_GATHER_FUSED):
 *   ES_GATHER_MATERIALIZE  projection per phoneme, then a row-gather kernel writes skip [B,T,dx2]
 *   ES_GATHER_FUSED        projection per phoneme; the first decoder block gathers rows of the table itself
 *                          ([B,T,dx2] is never written); decoders without that kernel variant (dx2 = 256, SIMT
 *                          mode) fall back to ES_GATHER_MATERIALIZE.
 * Both compute the same function (bit-identical).  (Value 0, a projection GEMM per frame, was removed.) */
#define ES_GATHER_MATERIALIZE 1
#define ES_GATHER_FUSED       2
int  es_model_set_decoder_gather(es_model_t* m, int mode);

/* Scratch requirement (bytes) of the calls below for a batch of B utterances, N phonemes,
 * T frames (T may be 0 for encoder-only use). */
size_t es_workspace_bytes(const es_model_t* m, int B, int N, int T);

/*
 * Replaces PhonemeEncoder.forward up to (not including) the feature upsampler
 * (networks.py:336-384): embedding, 2 encoder blocks, fuse, the three predictors, variance
 * embeddings, concat, duration rounding/clamp, plus the integer scan of the length regulator.
 *   phoneme       [B,N] int32
 *   phoneme_mask  [B,N] uint8 or NULL (the reference's B==1 mask-free path, networks.py:338)
 *   pitch_tgt / energy_tgt [B,N] f32, dur_tgt [B,N] int32: teacher-forcing targets
 *                 (train=True) or NULL (free running: predictions are embedded / rounded)
 * outputs
 *   pitch_pred, energy_pred, dur_pred  [B,N] f32   (the reference's [B,N,1])
 *   fused4        [B,N,4d] f32  concat [fused | pitch_emb | energy_emb | duration_feat]
 *   dur_int       [B,N] int32   durations driving the length regulator (networks.py:379-384,234)
 *   dur_cum       [B,N] int32   inclusive prefix sum of dur_int
 *   mel_len       [B]   int32   (networks.py:255)
 */
int es_encoder_forward(es_model_t* m, void* stream, int B, int N,
                       const int32_t* phoneme, const uint8_t* phoneme_mask,
                       const float* pitch_tgt, const float* energy_tgt, const int32_t* dur_tgt,
                       float* pitch_pred, float* energy_pred, float* dur_pred,
                       float* fused4, int32_t* dur_int, int32_t* dur_cum, int32_t* mel_len,
                       void* workspace, size_t workspace_bytes);

/*
 * Replaces FeatureUpsampler.forward (networks.py:228-258): materialises the expanded
 * features.  src[b,t] = min{n : dur_cum[b,n] > t} for t < mel_len[b], else -1.
 *   features [B,T,4d] f32 (zeros on padding), frame_mask [B,T] uint8 (1 = padding; the
 *   reference's [B,T,4d] bool mask is this broadcast over channels, ORed with the source
 *   phoneme's mask), src [B,T] int32 (may be NULL).
 */
int es_length_regulate(es_model_t* m, void* stream, int B, int N, int T,
                       const float* fused4, const int32_t* dur_cum, const uint8_t* phoneme_mask,
                       float* features, uint8_t* frame_mask, int32_t* src);

/*
 * The length regulator as the index map the gathered decoder entry uses internally
 * (FeatureUpsampler.forward, networks.py:228-258, as integers only):
 *   rows[b,t] = b*N + min{n : dur_cum[b,n] > t}  for t < mel_len[b],  B*N (the zero-padded row) otherwise.
 * rows [B,T] int32.  Exposed so that tests can check the map of the hot path bit-exactly.
 */
int es_frame_rows(es_model_t* m, void* stream, int B, int N, int T,
                  const int32_t* dur_cum, const int32_t* mel_len, int32_t* rows);

/*
 * Replaces LJSpeechDataModule.collate_fn (datamodule.py:29-76) and get_mask_from_lengths (utils/tools.py:43-51) for the
 * acoustic model's inputs -- the step BEFORE the path, on the device.  The B utterances arrive concatenated:
 * utterance u owns elements [offsets[u], offsets[u+1]) of the *_flat arrays (offsets [B+1] int32).  Outputs, all
 * device buffers: perm [B] (row r holds utterance perm[r]: STABLE argsort of decreasing length -- the reference's
 * np.argsort(-len) leaves the order of equal lengths unspecified), phoneme [B,N] int32 zero-padded, phoneme_mask [B,N]
 * (1 = padding), phoneme_len [B]; and, when the matching *_flat input is given (training batches), pitch / energy [B,N]
 * fp32, duration [B,N] int32 and mel_len [B] = sum of durations.  N >= the longest utterance (the caller knows the
 * lengths: it built the offsets).  B <= 12000.
 */
int es_collate(void* stream, int B, int N, const int32_t* offsets, const int32_t* phoneme_flat, const float* pitch_flat,
               const float* energy_flat, const int32_t* duration_flat, int32_t* perm, int32_t* phoneme,
               uint8_t* phoneme_mask, int32_t* phoneme_len, float* pitch, float* energy, int32_t* duration,
               int32_t* mel_len);

/* Opt-in, NOT part of the reference path: an fp16 copy (round to nearest even) of n fp32 values, for consumers that pull
 * the mel over PCIe (the fp32 mel [B,T,80] of configs[1] is 62.9 MB per step; the e2e of bench.py is bound by that
 * copy).  The fp32 mel remains the output of es_decoder_forward*; both buffers 16-byte aligned. */
int es_mel_to_half(void* stream, const float* mel, void* mel_f16, size_t n);

/* Replaces MelDecoder.forward (networks.py:291-304): features [B,T,4d] -> mel [B,T,n_mel]. */
int es_decoder_forward(es_model_t* m, void* stream, int B, int T,
                       const float* features, float* mel,
                       void* workspace, size_t workspace_bytes);

/*
 * Replaces the decoder half of Phoneme2Mel.forward (networks.py:422-427) without ever
 * materialising [B,T,4d].  MelDecoder.proj (networks.py:292) is row-wise and the length regulator
 * (networks.py:228-258) is a row gather, so they commute: the projection runs once per PHONEME
 * (B*N rows, + one row for the zero-padded frames), an integer kernel builds the frame -> row map,
 * and the first decoder block reads its input / skip rows through that map.  Frames t >= mel_len[b]
 * of the mel are zeroed when zero_padded_frames != 0 (the reference does so only when B > 1).
 * Workspace: es_workspace_bytes(m, B, N, T).
 */
int es_decoder_forward_gathered(es_model_t* m, void* stream, int B, int N, int T,
                                const float* fused4, const int32_t* dur_cum, const int32_t* mel_len,
                                int zero_padded_frames, float* mel,
                                void* workspace, size_t workspace_bytes);

/* Format of the split-fp16 tensor-core image (`*_w_h16`) of a dense phoneme-side layer with K input
 * channels, n_out output channels, `taps` conv taps and the given stride (the encoder blocks' qkv /
 * proj / ffn1 / ffn2 / merge weights, the predictors' conv1 / conv2):
 *   0  none: the layer runs on the fp32 SIMT kernel, the pointer is ignored
 *   1  resident image  [taps][2 (hi,lo)][K/8][n_out][8] halves           (packing.canon_split_taps)
 *   2  streamed units  [n_out/128][K/32][taps][2][4][128][8] halves      (packing.canon_split_units)
 *   3  streamed units  [n_out/256][K/32][taps][2][4][256][8] halves
 * The host packs what this returns; the library dispatches on the same rule. */
int es_dense_layout(int K, int n_out, int taps, int stride);

/* The tcgen05 kernels bound every mbarrier wait; a timeout sets a device flag instead of hanging
 * the GPU.  This call synchronises `stream` and returns non-zero if the flag was raised. */
int es_check_async_errors(void* stream);

/* Debug aid: when non-NULL, CTA 0 of every subsequent tcgen05 decoder launch writes clock64() stamps
 * [4 roles][32 tiles][8 events] (int64, device memory) -- tools/trace_decoder.py.  NULL turns it off. */
int es_debug_set_trace(void* dev_buf_i64);
/* Same for the fused phoneme kernel: [2 utterances][128] (clock64, event code) int64 pairs of CTA 0 -- one pair per
 * GEMM completion, weight wait, phase sync and row reduction (tools/trace_phoneme.py).  NULL turns it off. */
int es_debug_set_phoneme_trace(void* dev_buf_i64);
/* Number of kernels the library has launched since process start (bench.py's gpu_launches). */
uint64_t es_launch_count(void);

/* Per-kernel device timing for bench.py's roofline: between es_profile_begin and
 * es_profile_end every launch is bracketed by CUDA events on the launching stream.
 * es_profile_collect synchronises the recorded events and copies (kind, milliseconds)
 * pairs to HOST arrays; kinds are the ES_K_* values below. */
#define ES_K_EMBED      0
#define ES_K_ENC_GEMM   1
#define ES_K_ATTENTION  2
#define ES_K_FUSE       3
#define ES_K_PREDICTOR  4
#define ES_K_VARIANCE   5
#define ES_K_LENREG     6
#define ES_K_DEC_PROJ   7
#define ES_K_DEC_LAYER  8
#define ES_K_MEL        9
#define ES_K_POOLMASK   10
#define ES_K_PHONEME    11   /* whole phoneme side in one launch (es_umma_phoneme.cu) */
int es_profile_begin(int max_records);
int es_profile_end(void);
int es_profile_collect(int32_t* kinds_host, float* ms_host, int capacity, int* n_out);

/* Self-test of the tcgen05/TMA building block: C[M,N] = A[M,K] B[N,K]^T with split-fp16
 * operands (3 MMAs), M multiple of 128, N in {128,256}, K multiple of 64.  fp32 in/out. */
int es_selftest_umma_gemm(void* stream, int M, int N, int K, const float* A, const float* Bm, float* C);
/* The attention core of one encoder block in isolation (layers/blocks.py:44-63: softmax(scale q k^T) v over ALL n keys,
 * every head full width C): qkv [B,n,3*H*C] -> out [B,n,H*C].  tensor_core != 0: the tcgen05 kernels (C in {32, 64} with
 * n <= 128; C = 128 with n <= 128; C = 256 with n <= 64) -- returns 2 when the shape is outside their envelope;
 * tensor_core == 0: the fp32 SIMT kernel. */
int es_selftest_attention(void* stream, int B, int n, int C, int H, float scale, const float* qkv, float* out, int tensor_core);

/* =====================================================================================================================
 * HiFi-GAN generator -- the step AFTER the acoustic path.
 * Replaces hifigan.Generator.forward (hifigan/models.py:111-127; ResBlock1 :18-57) as called at model.py:161-162:
 *     wav = self.hifigan(mel.transpose(1, 2)).squeeze(1)
 * Weights are the plain conv weights after remove_weight_norm() (model.py:44), fp32, device, reference layouts:
 * Conv1d [Cout][Cin][K], ConvTranspose1d [Cin][Cout][K].  Only the ResBlock1 family ("resblock": "1", LJ_V2 and LJ_V1).
 */
#define ES_HG_MAX_UPS 6
#define ES_HG_MAX_RES 4

typedef struct es_hifigan_config {
    int32_t n_mel;                                   /* 80 */
    int32_t initial_channel;                         /* upsample_initial_channel (128 for V2) */
    int32_t n_up;                                    /* len(upsample_rates) */
    int32_t up_rate[ES_HG_MAX_UPS];                  /* [8, 8, 2, 2] */
    int32_t up_kernel[ES_HG_MAX_UPS];                /* [16, 16, 4, 4] */
    int32_t n_res;                                   /* len(resblock_kernel_sizes) */
    int32_t res_kernel[ES_HG_MAX_RES];               /* [3, 7, 11] */
    int32_t res_dilation[ES_HG_MAX_RES][3];          /* [[1, 3, 5]] * 3 */
} es_hifigan_config_t;

typedef struct es_hg_conv_w { const float* w; const float* b; } es_hg_conv_w_t;
typedef struct es_hg_resblock_w { es_hg_conv_w_t convs1[3]; es_hg_conv_w_t convs2[3]; } es_hg_resblock_w_t;
typedef struct es_hifigan_weights {
    es_hg_conv_w_t conv_pre;                                         /* [C0][n_mel][7] */
    es_hg_conv_w_t ups[ES_HG_MAX_UPS];                               /* [C_i][C_i / 2][k_i] */
    es_hg_resblock_w_t res[ES_HG_MAX_UPS][ES_HG_MAX_RES];            /* resblocks[i * n_res + j] */
    es_hg_conv_w_t conv_post;                                        /* [1][C_last][7] */
} es_hifigan_weights_t;

typedef struct es_hifigan es_hifigan_t;

int    es_hifigan_create(const es_hifigan_config_t* cfg, const es_hifigan_weights_t* w, es_hifigan_t** out);
void   es_hifigan_destroy(es_hifigan_t* h);
size_t es_hifigan_workspace_bytes(const es_hifigan_t* h, int B, int T);
/* mel: element (b, c, t) at mel[b * mel_sb + c * mel_sc + t * mel_st] -- [B,80,T] as the reference passes it
 * (sb = 80 T, sc = T, st = 1) or the acoustic model's own [B,T,80] (sb = 80 T, sc = 1, st = 80), so the transpose of
 * model.py:160 never materialises.  wav [B, T * prod(up_rate)] fp32 in (-1, 1). */
int    es_hifigan_forward(es_hifigan_t* h, void* stream, int B, int T, const float* mel, long long mel_sb, long long mel_sc,
                          long long mel_st, float* wav, void* workspace, size_t workspace_bytes);

/* =====================================================================================================================
 * Training-step pieces around the network.  The network's backward is not built; these are
 * the parts of EfficientSpeech.training_step that do not depend on it.
 */
/* Replaces EfficientSpeech.loss + the weighted total of training_step (model.py:167-217).  Predictions as the forward
 * returns them (mel [B,T,n_mel], pitch / energy / duration [B,N]); targets mel [B,T,n_mel], pitch / energy [B,N] fp32,
 * duration [B,N] int32; valid frames t < mel_len[b], valid phonemes phoneme_mask == 0 (null: all).  losses[5] (device) =
 * total (10 mel + 2 pitch + 2 energy + duration), mel L1, pitch MSE, energy MSE, log-duration MSE.  d_* (each optional):
 * gradient of the TOTAL with respect to the matching prediction.  Deterministic (fixed-order reductions). */
size_t es_loss_workspace_bytes(void);
int    es_loss(void* stream, int B, int N, int T, int n_mel, const float* mel_pred, const float* mel_tgt, const int32_t* mel_len,
               const float* pitch_pred, const float* energy_pred, const float* dur_pred, const float* pitch, const float* energy,
               const int32_t* duration, const uint8_t* phoneme_mask, float* losses, float* d_mel, float* d_pitch,
               float* d_energy, float* d_dur, void* workspace, size_t workspace_bytes);
/* Replaces one torch.optim.AdamW.step (model.py:279-283) over flat fp32 buffers of n elements.  The caller passes what
 * torch computes on the host in double: step_size = lr / (1 - beta1^t), bias_correction2_sqrt = sqrt(1 - beta2^t). */
int    es_adamw_step(void* stream, size_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float lr,
                     float beta1, float beta2, float eps, float weight_decay, float step_size, float bias_correction2_sqrt);

/* ---- differentiable primitives of the training step (csrc/es_train_ops.cu) --------------
 * Row-major fp32 activations [rows][C] ("B T C"), torch-layout parameters.  Together they stand in for what
 * lightning's training_step runs through torch autograd on layers/networks.py (model.py:97-113): every call is one
 * operator or its adjoint; efficientspeech_b200/train_ops.py composes them.  act kinds: 0 none, 1 ReLU, 2 GELU (erf), 3 tanh. */
/* C[b] (+)= op(A[b]) op(B[b]) (+ bias[n]);  op(A) is M x K, op(B) is K x N; trans_x: the stored matrix is the transpose.
 * k_chunk > 0: split-K -- slice b of `batch` takes k in [b k_chunk, (b+1) k_chunk) of ONE product and writes its partial
 * to C + b stride_c (weight gradients contract over every frame of the batch; es_t_colsum adds the partials in order). */
int es_t_gemm(void* stream, int batch, int M, int N, int K, const float* A, int lda, long long stride_a, int trans_a,
              const float* B, int ldb, long long stride_b, int trans_b, float* C, int ldc, long long stride_c,
              const float* bias, int accumulate, int k_chunk, int grad_mask);
/* Products with N >= 16, K >= 16, M >= 32 (one matrix, or split-K) run on the tensor cores: each fp32 operand as two
 * 16-bit halves, three tcgen05 MMAs per product, fp32 accumulation (csrc/es_train_gemm.cu).  grad_mask bit 0 / 1: A / B
 * holds GRADIENTS; such a product splits both operands as bf16 pairs (fp32's exponent range) instead of fp16 pairs.  Everything else, and
 * everything when switched off here, runs the fp32 SIMT kernel. */
void es_t_set_tensor_core(int enable);
/* cols[b, t, c*k + tau] = X[b, t*s + tau - p, c] (zero outside; a row is ordered like torch's flattened Conv1d weight),
 * and its adjoint (also ConvTranspose1d's forward scatter). */
int es_t_im2col(void* stream, const float* X, float* cols, int B, int n_in, int n_out, int C, int k, int s, int p);
int es_t_col2im(void* stream, const float* cols, float* X, int B, int n_in, int n_out, int C, int k, int s, int p, int accumulate);
/* depthwise Conv1d over time, "same" padding k/2 (layers/networks.py:281); w [C][k] */
int es_t_dwconv_fwd(void* stream, const float* X, const float* w, const float* bias, float* Y, int B, int T, int C, int k);
/* the weight / bias gradients reduce over every frame: row slices into `ws`, added in a fixed order (deterministic) */
size_t es_t_dwconv_bwd_workspace_floats(int B, int T, int C, int k);
int es_t_dwconv_bwd(void* stream, const float* dY, const float* X, const float* w, float* dX, float* dw, float* db, int B, int T, int C, int k,
                    float* ws, size_t ws_floats);
int es_t_layernorm_fwd(void* stream, const float* X, const float* g, const float* b, float* Y, float* xhat, float* rstd, long long rows, int C);
/* one pass: dX and, per block, partial sums of dg / db in `ws`, then added in a fixed order.  C <= 256. */
size_t es_t_layernorm_bwd_workspace_floats(long long rows, int C);
int es_t_layernorm_bwd(void* stream, const float* dY, const float* xhat, const float* rstd, const float* g, float* dX, float* dg, float* db,
                       long long rows, int C, float* ws, size_t ws_floats);
/* out[c] (+)= sum_r A[r,c] (* B[r,c] when B is given); long reductions go through row slices in `ws` (may be null when
 * es_t_colsum_workspace_floats returns 0) */
size_t es_t_colsum_workspace_floats(long long rows, int C);
int es_t_colsum(void* stream, const float* A, const float* B, float* out, long long rows, int C, int accumulate, float* ws, size_t ws_floats);
int es_t_act_fwd(void* stream, const float* X, float* Y, long long n, int kind);
/* saved: the OUTPUT for ReLU / tanh, the INPUT for GELU */
int es_t_act_bwd(void* stream, const float* dY, const float* saved, float* dX, long long n, int kind);
int es_t_softmax_fwd(void* stream, const float* X, float* Y, long long rows, int n, float scale);
int es_t_softmax_bwd(void* stream, const float* dY, const float* Y, float* dX, long long rows, int n, float scale);
/* out[r] = table[idx[r]] (idx < 0: zeros); d_table[idx[r]] += d_out[r] except rows idx == skip_index (padding_idx) */
int es_t_gather_rows(void* stream, const float* table, const int32_t* idx, float* out, long long rows, int C);
int es_t_scatter_add_rows(void* stream, const float* d_out, const int32_t* idx, float* d_table, long long rows, int C, int skip_index);
/* length regulator (layers/blocks.py LengthRegulator): cum = inclusive cumulative durations [B,N]; frames past the sum
 * are zero.  reduce_rows is its adjoint, a fixed-order segmented sum. */
int es_t_expand_rows(void* stream, const float* in, const int32_t* cum, float* out, int B, int N, int T, int C);
int es_t_reduce_rows(void* stream, const float* d_out, const int32_t* cum, float* d_in, int B, int N, int T, int C);
/* torch.bucketize(v, bins): first i with bins[i] >= v (layers/networks.py:128-141 on the TARGET pitch / energy) */
int es_t_bucketize(void* stream, const float* v, const float* bins, int n_bins, int32_t* out, long long n);
int es_t_axpby(void* stream, const float* X, const float* Y, float* out, long long n, float a, float b);
int es_t_mask_rows(void* stream, const float* X, const uint8_t* mask, float* Y, long long rows, int C);
int es_t_copy2d(void* stream, const float* src, int lds, float* dst, int ldd, long long rows, int cols, int accumulate);

#ifdef __cplusplus
}
#endif
#endif /* ES_B200_H */

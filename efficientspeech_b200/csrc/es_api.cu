// C ABI (include/es_b200.h): model handle, workspace planning and the kernel sequence of the
// acoustic forward path.  Host code only; every kernel lives in the sibling .cu files.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "es_common.cuh"
#include "es_kernels.cuh"

namespace es {

static thread_local std::string g_error;
std::atomic<uint64_t> g_launches{0};
void set_error(const std::string& msg) { g_error = msg; }

// per-launch event timing -------------------------------------------------------------------
namespace {
struct Profiler {
    bool on = false;
    size_t cap = 0;
    std::vector<cudaEvent_t> ev0, ev1;
    std::vector<int> kinds;
    size_t n = 0;
    bool open = false;
} g_prof;
}  // namespace

void prof_begin_range(int kind, cudaStream_t s) {
    if (!g_prof.on || g_prof.n >= g_prof.cap) return;
    g_prof.kinds[g_prof.n] = kind;
    cudaEventRecord(g_prof.ev0[g_prof.n], s);
    g_prof.open = true;
}
void prof_end_range(cudaStream_t s) {
    if (!g_prof.on || !g_prof.open) return;
    cudaEventRecord(g_prof.ev1[g_prof.n], s);
    g_prof.n++;
    g_prof.open = false;
}

}  // namespace es

struct es_model {
    es_config_t cfg;
    es_weights_t w;
    int use_tensor_core;
    int gather_mode;        // es_model_set_decoder_gather: ES_GATHER_* (how the length regulator meets the decoder)
    int fused_phoneme;      // es_model_set_fused_phoneme: whole phoneme side in one kernel where supported
    int ragged_schedule;    // es_model_set_ragged_schedule: the decoder skips tiles that cannot reach a valid frame
    // derived geometry
    int d, C[2], H[2], k[2], hC[2], dx4, dx2, n_layers;
};

namespace {

using namespace es;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over the caller's workspace; with base == nullptr it only measures.
struct Arena {
    char* base;
    size_t off = 0;
    explicit Arena(void* b) : base(static_cast<char*>(b)) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

inline int enc_n1(const es_model* m, int N) {
    const int k1 = m->k[1];
    return (N + 2 * (k1 / 2) - k1) / 2 + 1;           // Conv1d(stride 2) output length (networks.py:40)
}

struct EncBufs {
    float *x0, *qkv, *att, *x1, *h, *feat0, *xm1, *feat1, *fused, *y1[3], *dur_feat;
    uint8_t* mask1;
};

EncBufs plan_encoder(const es_model* m, Arena& a, int B, int N) {
    EncBufs e;
    const size_t n1 = enc_n1(m, N);
    const size_t r0 = (size_t)B * N, r1 = (size_t)B * n1;
    auto mx = [](size_t x, size_t y) { return x > y ? x : y; };
    e.x0 = a.take<float>(r0 * m->C[0]);
    e.qkv = a.take<float>(mx(r0 * 3 * m->H[0] * m->C[0], r1 * 3 * m->H[1] * m->C[1]));
    e.att = a.take<float>(mx(r0 * m->H[0] * m->C[0], r1 * m->H[1] * m->C[1]));
    e.x1 = a.take<float>(mx(r0 * m->C[0], r1 * m->C[1]));
    e.h = a.take<float>(mx(r0 * m->hC[0], r1 * m->hC[1]));
    e.feat0 = a.take<float>(r0 * m->C[0]);
    e.xm1 = a.take<float>(r1 * m->C[1]);
    e.feat1 = a.take<float>(r1 * m->C[1]);
    e.fused = a.take<float>(r0 * m->d);
    for (int i = 0; i < 3; ++i) e.y1[i] = a.take<float>(r0 * m->d);
    e.dur_feat = a.take<float>(r0 * m->d);
    e.mask1 = a.take<uint8_t>(r1);
    return e;
}

struct DecBufs {
    float* buf[3];
    float* P;            // gathered entry (N > 0): per-phoneme projection table [B*N + 1][dx2], last row = padded frames
    int* src;            //                         frame -> table row map [B*T]
    int pad_id;          //                         = B*N, the table row of the zero-padded frames
    int2* tiles;         // ragged schedule (es_gather.cu): (b, t0) of the tiles that can reach a valid frame, [B * ceil(T/64)]
    int* tile_count;     //                                 their number
};

DecBufs plan_decoder(const es_model* m, Arena& a, int B, int N, int T) {
    DecBufs d;
    for (int i = 0; i < 3; ++i) d.buf[i] = a.take<float>((size_t)B * T * m->dx2);
    d.P = nullptr; d.src = nullptr; d.pad_id = B * N; d.tiles = nullptr; d.tile_count = nullptr;
    if (N > 0) {
        d.P = a.take<float>(((size_t)B * N + 1) * m->dx2);
        d.src = a.take<int>((size_t)B * T);
        d.tiles = a.take<int2>((size_t)B * ((T + 63) / 64));
        d.tile_count = a.take<int>(1);
    }
    return d;
}

size_t decoder_workspace_bytes(const es_model* m, int B, int N, int T) {
    Arena a(nullptr);
    plan_decoder(m, a, B, N, T);
    return align_up(a.off, 256) + 256;
}

RowGemmParams base_params(int B, int n_in, int n_out, int K, int Nout, const float* A, int lda,
                          const float* W, float* Y, int ldy) {
    RowGemmParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.n_in = n_in; p.n_out = n_out; p.K = K; p.Nout = Nout;
    p.ldw = (Nout + 31) / 32 * 32;
    p.taps = 1; p.stride = 1; p.pad = 0; p.mode = ROW_PLAIN;
    p.A = A; p.lda = lda; p.W = W; p.Y = Y; p.ldy = ldy;
    return p;
}

// Dense layer: tcgen05 row GEMM when the layer is inside its envelope, fp32 SIMT otherwise.
int gemm(const es_model* m, const RowGemmParams& p, const void* w_h16, cudaStream_t s) {
    if (m->use_tensor_core && w_h16) {
        const int lay = dense_layout(p.K, p.Nout, p.taps, p.stride);     // decides the packed image's format too
        const int rc = lay == 1 ? launch_umma_rowgemm(p, w_h16, s) : lay >= 2 ? launch_umma_wide(p, w_h16, s) : -1;
        if (rc >= 0) return rc;
    }
    return launch_rowgemm(p, s);
}

// One encoder block after its merge conv (networks.py:72-85).
int encoder_block(const es_model* m, int i, int B, int n, const float* x_in, const uint8_t* mask,
                  const EncBufs& e, float* feat_out, cudaStream_t s) {
    const es_enc_block_w_t& w = m->w.enc[i];
    const int C = m->C[i], H = m->H[i], hC = m->hC[i];
    // qkv = x Wqkv^T                                                            blocks.py:45
    RowGemmParams p = base_params(B, n, n, C, 3 * H * C, x_in, C, w.qkv_w, e.qkv, 3 * H * C);
    { ProfRange r(ES_K_ENC_GEMM, s); if (gemm(m, p, w.qkv_w_h16, s)) return 1; }
    // softmax(QK^T scale) V, all keys (mask never applied)                         blocks.py:49-65
    const float scale = 1.0f / sqrtf((float)(C / H));
    {
        ProfRange r(ES_K_ATTENTION, s);
        int rc = m->use_tensor_core ? launch_umma_attention(e.qkv, e.att, B, n, C, H, scale, s) : -1;
        if (rc > 0) return 1;
        if (rc < 0 && launch_attention(e.qkv, e.att, B, n, C, H, scale, s)) return 1;
    }
    // x1 = mask(LN1(proj(att) + x))                                                blocks.py:66, networks.py:73-75
    p = base_params(B, n, n, H * C, C, e.att, H * C, w.proj_w, e.x1, C);
    p.bias = w.proj_b; p.res1 = x_in; p.ldr1 = C; p.ln_g = w.ln1_g; p.ln_b = w.ln1_b; p.row_mask = mask;
    { ProfRange r(ES_K_ENC_GEMM, s); if (gemm(m, p, w.proj_w_h16, s)) return 1; }
    // h = GELU(conv3(mlp1(x1)))  with mlp1 folded into the conv taps               blocks.py:23-27
    p = base_params(B, n, n, C, hC, e.x1, C, w.ffn1_w, e.h, hC);
    p.taps = 3; p.pad = 1; p.bias = w.ffn1_b; p.tap_bias = w.ffn1_tapb; p.act1 = ACT_GELU;
    { ProfRange r(ES_K_ENC_GEMM, s); if (gemm(m, p, w.ffn1_w_h16, s)) return 1; }
    // feat = mask(LN2(mlp2(h) + x1))                                               blocks.py:28, networks.py:80-83
    p = base_params(B, n, n, hC, C, e.h, hC, w.ffn2_w, feat_out, C);
    p.bias = w.ffn2_b; p.res1 = e.x1; p.ldr1 = C; p.ln_g = w.ln2_g; p.ln_b = w.ln2_b; p.row_mask = mask;
    ProfRange r(ES_K_ENC_GEMM, s);
    return gemm(m, p, w.ffn2_w_h16, s);
}

// AcousticDecoder.forward (networks.py:151-165), stage 1: y1 = ReLU(LN1(ReLU(conv1(fused))))
RowGemmParams predictor_stage1(const es_model* m, const es_predictor_w_t& w, int B, int N, const float* fused, float* y1) {
    const int d = m->d;
    RowGemmParams p = base_params(B, N, N, d, d, fused, d, w.conv1_w, y1, d);
    p.taps = 3; p.pad = 1; p.bias = w.conv1_b; p.act1 = ACT_RELU;
    p.ln_g = w.ln1_g; p.ln_b = w.ln1_b; p.act2 = ACT_RELU;
    return p;
}
// stage 2: y = ReLU(conv2(y1)); pred = (ReLU)(y . w + b); duration only: feat = LN2(y)
RowGemmParams predictor_stage2(const es_model* m, const es_predictor_w_t& w, int B, int N, const float* y1, float* pred,
                               bool is_duration, float* feat_out) {
    const int d = m->d;
    RowGemmParams p = base_params(B, N, N, d, d, y1, d, w.conv2_w, is_duration ? feat_out : nullptr, d);
    p.taps = 3; p.pad = 1; p.bias = w.conv2_b; p.act1 = ACT_RELU;
    p.dot_w = w.lin_w; p.dot_b = w.lin_b; p.dot_out = pred; p.dot_relu = is_duration ? 1 : 0;
    if (is_duration) { p.ln_g = w.ln2_g; p.ln_b = w.ln2_b; }   // norm2 output is only consumed for duration
    return p;
}

// The three predictors are independent: each stage runs as ONE batched launch when the narrow kernel
// applies (d <= 96), else as three launches.
int predictors(const es_model* m, int B, int N, const float* fused, float* const y1[3], float* pitch_pred,
               float* energy_pred, float* dur_pred, float* dur_feat, cudaStream_t s) {
    const es_predictor_w_t* w[3] = {&m->w.pitch, &m->w.energy, &m->w.duration};
    float* preds[3] = {pitch_pred, energy_pred, dur_pred};
    RowGemmParams st1[3], st2[3];
    for (int i = 0; i < 3; ++i) {
        st1[i] = predictor_stage1(m, *w[i], B, N, fused, y1[i]);
        st2[i] = predictor_stage2(m, *w[i], B, N, y1[i], preds[i], i == 2, i == 2 ? dur_feat : nullptr);
    }
    // tensor-core path: the three predictors share their geometry, so each stage is ONE batched launch
    // (consecutive CTAs serve different predictors, each keeps its own weights resident)
    if (m->use_tensor_core && w[0]->conv1_w_h16 && w[0]->conv2_w_h16) {
        const void* w1[3] = {w[0]->conv1_w_h16, w[1]->conv1_w_h16, w[2]->conv1_w_h16};
        const void* w2[3] = {w[0]->conv2_w_h16, w[1]->conv2_w_h16, w[2]->conv2_w_h16};
        int rc;
        if (dense_layout(st1[0].K, st1[0].Nout, st1[0].taps, st1[0].stride) >= 2) {
            // wide predictors (base): streamed-weight kernel, one launch per predictor and stage
            for (int i = 0; i < 3; ++i) {
                ProfRange r(ES_K_PREDICTOR, s);
                rc = launch_umma_wide(st1[i], w1[i], s);
                if (rc > 0) return 1;
                if (rc < 0 && launch_rowgemm(st1[i], s)) return 1;
            }
            for (int i = 0; i < 3; ++i) {
                ProfRange r(ES_K_PREDICTOR, s);
                rc = launch_umma_wide(st2[i], w2[i], s);
                if (rc > 0) return 1;
                if (rc < 0 && launch_rowgemm(st2[i], s)) return 1;
            }
            return 0;
        }
        { ProfRange r(ES_K_PREDICTOR, s); rc = launch_umma_rowgemm_batch(st1, w1, 3, s); }
        if (rc > 0) return 1;
        if (rc == 0) {
            { ProfRange r(ES_K_PREDICTOR, s); rc = launch_umma_rowgemm_batch(st2, w2, 3, s); }
            if (rc > 0) return 1;
            if (rc < 0) for (int i = 0; i < 3; ++i) if (launch_rowgemm(st2[i], s)) return 1;
            return 0;
        }
    }
    // stage 2 differs between predictors only in pointers / LN2 / relu flags; geometry is identical
    {
        ProfRange r(ES_K_PREDICTOR, s);
        const int rc = launch_rowgemm_narrow_batch(st1, 3, s);
        if (rc > 0) return 1;
        if (rc < 0) for (int i = 0; i < 3; ++i) if (launch_rowgemm(st1[i], s)) return 1;
    }
    {
        ProfRange r(ES_K_PREDICTOR, s);
        const int rc = launch_rowgemm_narrow_batch(st2, 3, s);
        if (rc > 0) return 1;
        if (rc < 0) for (int i = 0; i < 3; ++i) if (launch_rowgemm(st2[i], s)) return 1;
    }
    return 0;
}

int decoder_layers(const es_model* m, int B, int T, DecBufs& db, int s_idx, const int* zero_from,
                   float* mel, cudaStream_t s, bool ragged = false);
bool decoder_all_umma256(const es_model* m);
int project_rows(const es_model* m, int rows, const float* in, float* out, cudaStream_t s);
bool decoder_all_umma128(const es_model* m);

}  // namespace

extern "C" {

int es_abi_version(void) { return ES_ABI_VERSION; }

int es_selftest_attention(void* stream, int B, int n, int C, int H, float scale, const float* qkv, float* out, int tensor_core) {
    ES_CHECK(qkv && out && B >= 1 && n >= 1 && C >= 1 && H >= 1, "bad arguments");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!tensor_core) return es::launch_attention(qkv, out, B, n, C, H, scale, s);
    const int rc = es::launch_umma_attention(qkv, out, B, n, C, H, scale, s);
    if (rc < 0) { es::set_error("es_selftest_attention: shape outside the tcgen05 attention envelope"); return 2; }
    return rc;
}
const char* es_last_error(void) { return es::g_error.c_str(); }
uint64_t es_launch_count(void) { return es::g_launches.load(); }
int es_debug_set_trace(void* dev_buf_i64) { es::umma_dec_set_trace(static_cast<long long*>(dev_buf_i64)); return 0; }
int es_debug_set_phoneme_trace(void* dev_buf_i64) { es::umma_phoneme_set_trace(static_cast<long long*>(dev_buf_i64)); return 0; }
int es_dense_layout(int K, int n_out, int taps, int stride) { return es::dense_layout(K, n_out, taps, stride); }
int es_check_async_errors(void* stream) { return es::umma_dec_check_errors(static_cast<cudaStream_t>(stream)); }

int es_profile_begin(int max_records) {
    ES_CHECK(max_records > 0 && max_records <= (1 << 20), "bad record count");
    auto& P = es::g_prof;
    P.on = false;
    for (size_t i = P.ev0.size(); i < (size_t)max_records; ++i) {
        cudaEvent_t a, b;
        ES_CUDA(cudaEventCreate(&a));
        ES_CUDA(cudaEventCreate(&b));
        P.ev0.push_back(a);
        P.ev1.push_back(b);
    }
    P.kinds.assign(P.ev0.size(), 0);
    P.cap = (size_t)max_records;
    P.n = 0;
    P.open = false;
    P.on = true;
    return 0;
}

int es_profile_end(void) {
    es::g_prof.on = false;
    return 0;
}

int es_profile_collect(int32_t* kinds_host, float* ms_host, int capacity, int* n_out) {
    ES_CHECK(kinds_host && ms_host && n_out, "null argument");
    auto& P = es::g_prof;
    int n = (int)P.n < capacity ? (int)P.n : capacity;
    for (int i = 0; i < n; ++i) {
        ES_CUDA(cudaEventSynchronize(P.ev1[i]));
        float ms = 0.f;
        ES_CUDA(cudaEventElapsedTime(&ms, P.ev0[i], P.ev1[i]));
        kinds_host[i] = P.kinds[i];
        ms_host[i] = ms;
    }
    *n_out = n;
    return 0;
}

int es_model_create(const es_config_t* cfg, const es_weights_t* w, es_model_t** out) {
    ES_CHECK(cfg && w && out, "null argument");
    ES_CHECK(cfg->dim % 32 == 0 && cfg->dim >= 32 && cfg->dim <= 128, "dim must be 32..128, multiple of 32");
    ES_CHECK(cfg->kernel_size == 3 || cfg->kernel_size == 5, "kernel_size must be 3 or 5");
    ES_CHECK(cfg->n_blocks >= 1 && cfg->n_blocks <= ES_MAX_DEC_BLOCKS, "n_blocks out of range");
    ES_CHECK(cfg->n_blocks * cfg->block_depth <= ES_MAX_DEC_LAYERS && cfg->block_depth >= 1, "too many decoder layers");
    ES_CHECK(cfg->decoder_kernel_size % 2 == 1 && cfg->decoder_kernel_size <= ES_MAX_TAPS, "bad decoder kernel size");
    ES_CHECK(cfg->n_mel >= 1 && cfg->n_mel <= 256, "n_mel out of range");
    ES_CHECK(cfg->head >= 1 && cfg->expansion >= 1, "bad head/expansion");
    es_model* m = new (std::nothrow) es_model;
    ES_CHECK(m, "out of memory");
    m->cfg = *cfg;
    m->w = *w;
    m->use_tensor_core = 1;
    m->gather_mode = ES_GATHER_FUSED;
    m->fused_phoneme = 1;
    m->ragged_schedule = 1;
    m->d = cfg->dim;
    m->C[0] = cfg->dim; m->C[1] = 2 * cfg->dim;
    m->H[0] = cfg->head; m->H[1] = 2 * cfg->head;
    m->k[0] = cfg->kernel_size; m->k[1] = cfg->kernel_size - 2;
    m->hC[0] = m->C[0] * cfg->expansion; m->hC[1] = m->C[1] * cfg->expansion;
    m->dx4 = 4 * cfg->dim;
    m->dx2 = m->dx4 < 256 ? m->dx4 : 256;
    m->n_layers = cfg->n_blocks * cfg->block_depth;
    *out = m;
    return 0;
}

void es_model_destroy(es_model_t* m) { delete m; }

int es_model_set_tensor_core(es_model_t* m, int enable) {
    ES_CHECK(m, "null model");
    m->use_tensor_core = enable ? 1 : 0;
    return 0;
}

int es_model_set_fused_phoneme(es_model_t* m, int enable) {
    ES_CHECK(m, "null model");
    m->fused_phoneme = enable ? 1 : 0;
    return 0;
}

int es_model_set_ragged_schedule(es_model_t* m, int enable) {
    ES_CHECK(m, "null model");
    m->ragged_schedule = enable ? 1 : 0;
    return 0;
}

int es_model_set_decoder_gather(es_model_t* m, int mode) {
    ES_CHECK(m, "null model");
    ES_CHECK(mode == ES_GATHER_MATERIALIZE || mode == ES_GATHER_FUSED, "unknown gather mode");
    m->gather_mode = mode;
    return 0;
}

size_t es_workspace_bytes(const es_model_t* m, int B, int N, int T) {
    if (!m || B <= 0) return 0;
    Arena a(nullptr);
    if (N > 0) plan_encoder(m, a, B, N);
    if (T > 0) plan_decoder(m, a, B, N, T);
    return align_up(a.off, 256) + 256;
}

int es_encoder_forward(es_model_t* m, void* stream, int B, int N,
                       const int32_t* phoneme, const uint8_t* phoneme_mask,
                       const float* pitch_tgt, const float* energy_tgt, const int32_t* dur_tgt,
                       float* pitch_pred, float* energy_pred, float* dur_pred,
                       float* fused4, int32_t* dur_int, int32_t* dur_cum, int32_t* mel_len,
                       void* workspace, size_t workspace_bytes) {
    ES_CHECK(m, "null model");
    ES_CHECK(B >= 1 && B <= 65535 && N >= 2, "need 1 <= B <= 65535 and N >= 2 phonemes");
    ES_CHECK(phoneme && pitch_pred && energy_pred && dur_pred && fused4 && dur_int && dur_cum && mel_len,
             "null tensor");
    ES_CHECK(workspace && workspace_bytes >= es_workspace_bytes(m, B, N, 0), "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Arena a(reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(workspace), 256)));
    EncBufs e = plan_encoder(m, a, B, N);
    const int d = m->d, n1 = enc_n1(m, N);

    // tiny geometry, N <= 128: the whole phoneme side in ONE kernel, activations never leave the SM (es_umma_phoneme.cu)
    if (m->use_tensor_core && m->fused_phoneme && umma_phoneme_supported(m->cfg, m->w, N) && n1 <= 64) {
        const int pool = (int)nearbyintf((float)((double)N / (double)n1));
        int rc;
        { ProfRange r(ES_K_PHONEME, s);
          rc = launch_umma_phoneme(m->cfg, m->w, B, N, n1, pool, phoneme, phoneme_mask, pitch_tgt, energy_tgt, dur_tgt,
                                   pitch_pred, energy_pred, dur_pred, fused4, dur_int, dur_cum, mel_len, e.xm1, s); }
        if (rc > 0) return 1;
        if (rc == 0) return 0;
    }
    // block 0: embedding + merge conv + 1x1 as k table gathers (networks.py:54,64-67)
    { ProfRange r(ES_K_EMBED, s); if (launch_embed_merge(phoneme, m->w.enc[0].merge_w, e.x0, B, N, m->C[0], m->k[0], m->cfg.n_symbols, s)) return 1; }
    if (encoder_block(m, 0, B, N, e.x0, phoneme_mask, e, e.feat0, s)) return 1;
    // block 1 merge: Conv1d(d,d,k-2,stride 2) . Conv1d(d,2d,1) folded (networks.py:64-67)
    {
        RowGemmParams p = base_params(B, N, n1, m->C[0], m->C[1], e.feat0, m->C[0], m->w.enc[1].merge_w, e.xm1, m->C[1]);
        p.taps = m->k[1]; p.stride = 2; p.pad = m->k[1] / 2;
        const void* img = m->w.enc[1].merge_w_h16;
        if (m->use_tensor_core && m->w.enc[1].merge2_w_h16 && m->k[1] == 3 && N % 2 == 0 && n1 == N / 2) {
            // paired rows: [N][C] read as [N/2][2C], stride-1 conv with K = 2C (es_b200.h: merge2_w)
            p = base_params(B, N / 2, n1, 2 * m->C[0], m->C[1], e.feat0, 2 * m->C[0], m->w.enc[1].merge2_w, e.xm1, m->C[1]);
            p.taps = 3; p.stride = 1; p.pad = 1;
            img = m->w.enc[1].merge2_w_h16;
        }
        ProfRange r(ES_K_ENC_GEMM, s);
        if (gemm(m, p, img, s)) return 1;
    }
    const uint8_t* mask1 = nullptr;
    if (phoneme_mask) {
        // pool = round(N / n1), torch.round of an fp32 tensor: half-to-even (networks.py:69-70)
        const int pool = (int)nearbyintf((float)((double)N / (double)n1));
        { ProfRange r(ES_K_POOLMASK, s); if (launch_pool_mask(phoneme_mask, e.mask1, B, N, n1, pool, s)) return 1; }
        mask1 = e.mask1;
    }
    if (encoder_block(m, 1, B, n1, e.xm1, mask1, e, e.feat1, s)) return 1;
    // fuse (networks.py:189-219)
    bool fused_done = false;
    if (m->use_tensor_core && m->w.fuse_u_h16 && m->w.fuse_a0_h16) {
        // tensor-core form: U = feat1 [G_0|..|G_{k-1}] + [g_0|..|g_{k-1}] for every half-rate position (e.qkv is
        // free by now), then fused = mask(c + A0 feat0 + the stride-2 scatter of U) in the second GEMM's epilogue.
        // d % 128 == 0: streamed-weight kernel (es_umma_wide.cu); d <= 64: resident-weight row GEMM (es_umma_enc.cu).
        // The packed images follow the same rule (modules.py).
        const int k = m->k[0];
        const bool wide = d % 128 == 0;
        const bool narrow = !wide && dense_layout(2 * d, k * d, 1, 1) == 1 && dense_layout(d, d, 1, 1) == 1;
        float* U = e.qkv;
        if (wide || narrow) {
            RowGemmParams p = base_params(B, n1, n1, 2 * d, k * d, e.feat1, 2 * d, nullptr, U, k * d);
            p.bias = m->w.fuse_gb;
            int rc;
            { ProfRange r(ES_K_FUSE, s); rc = wide ? launch_umma_wide(p, m->w.fuse_u_h16, s, 128) : launch_umma_rowgemm(p, m->w.fuse_u_h16, s); }
            if (rc > 0) return 1;
            if (rc == 0) {
                p = base_params(B, N, N, d, d, e.feat0, d, nullptr, e.fused, d);
                p.bias = m->w.fuse_c; p.row_mask = phoneme_mask;
                p.fuse_u = U; p.fuse_k = k; p.fuse_n1 = n1; p.fuse_ld = k * d;
                ProfRange r(ES_K_FUSE, s);
                rc = wide ? launch_umma_wide(p, m->w.fuse_a0_h16, s, 128) : launch_umma_rowgemm(p, m->w.fuse_a0_h16, s);
                if (rc != 0) { ES_CHECK(rc < 0, "fuse GEMM failed"); ES_CHECK(false, "fuse epilogue outside the tensor-core kernels' envelope"); }
                fused_done = true;
            }
        }
    }
    if (!fused_done) { ProfRange r(ES_K_FUSE, s);
      if (launch_fuse(e.feat0, e.feat1, m->w.fuse_a0, m->w.fuse_g, m->w.fuse_gb, m->w.fuse_c, phoneme_mask,
                      e.fused, B, N, n1, d, m->k[0], s)) return 1; }
    // predictors (networks.py:349,357,366)
    if (predictors(m, B, N, e.fused, e.y1, pitch_pred, energy_pred, dur_pred, e.dur_feat, s)) return 1;
    // variance embeddings, concat, duration rounding, integer scan (networks.py:349-384, 234, 255)
    ProfRange r(ES_K_VARIANCE, s);
    return launch_variance_scan(e.fused, e.dur_feat, pitch_pred, energy_pred, dur_pred, pitch_tgt, energy_tgt,
                                dur_tgt, phoneme_mask, m->w.pitch, m->w.energy, fused4, dur_int, dur_cum,
                                mel_len, B, N, d, s);
}

int es_length_regulate(es_model_t* m, void* stream, int B, int N, int T,
                       const float* fused4, const int32_t* dur_cum, const uint8_t* phoneme_mask,
                       float* features, uint8_t* frame_mask, int32_t* src) {
    ES_CHECK(m, "null model");
    ES_CHECK(B >= 1 && N >= 1 && T >= 0, "bad shape");
    ES_CHECK(dur_cum && (fused4 || !features), "null tensor");
    ProfRange r(ES_K_LENREG, static_cast<cudaStream_t>(stream));
    return launch_length_regulate(fused4, dur_cum, phoneme_mask, features, frame_mask, src, B, N, T, m->dx4,
                                  static_cast<cudaStream_t>(stream));
}

int es_frame_rows(es_model_t* m, void* stream, int B, int N, int T,
                  const int32_t* dur_cum, const int32_t* mel_len, int32_t* rows) {
    ES_CHECK(m, "null model");
    ES_CHECK(dur_cum && mel_len && rows, "null tensor");
    ProfRange r(ES_K_LENREG, static_cast<cudaStream_t>(stream));
    return launch_frame_source(dur_cum, mel_len, rows, B, N, T, nullptr, nullptr, nullptr, 0, nullptr,
                               static_cast<cudaStream_t>(stream));
}

int es_collate(void* stream, int B, int N, const int32_t* offsets, const int32_t* phoneme_flat, const float* pitch_flat,
               const float* energy_flat, const int32_t* duration_flat, int32_t* perm, int32_t* phoneme,
               uint8_t* phoneme_mask, int32_t* phoneme_len, float* pitch, float* energy, int32_t* duration,
               int32_t* mel_len) {
    return launch_collate(B, N, offsets, phoneme_flat, pitch_flat, energy_flat, duration_flat, perm, phoneme, phoneme_mask,
                          phoneme_len, pitch, energy, duration, mel_len, static_cast<cudaStream_t>(stream));
}

int es_mel_to_half(void* stream, const float* mel, void* mel_f16, size_t n) {
    return launch_cast_f32_f16(mel, mel_f16, n, static_cast<cudaStream_t>(stream));
}

int es_decoder_forward(es_model_t* m, void* stream, int B, int T, const float* features, float* mel,
                       void* workspace, size_t workspace_bytes) {
    ES_CHECK(m, "null model");
    ES_CHECK(B >= 1 && B <= 65535 && T >= 1, "need 1 <= B <= 65535 and T >= 1 frames");
    ES_CHECK(features && mel, "null tensor");
    ES_CHECK(workspace && workspace_bytes >= decoder_workspace_bytes(m, B, 0, T), "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Arena a(reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(workspace), 256)));
    DecBufs db = plan_decoder(m, a, B, 0, T);
    // skip = LN(tanh(Linear(features)))                                           networks.py:292
    if (m->use_tensor_core && m->w.dproj_w_h16 && m->dx4 == 128 &&
        umma_dec_supported(m->dx2, m->cfg.decoder_kernel_size, m->dx2)) {
        { ProfRange r(ES_K_DEC_PROJ, s);
          if (launch_umma_dec(2, B, T, m->dx2, features, nullptr, nullptr, m->w.dproj_w_h16,
                              m->w.dproj_b, 1, m->w.dproj_ln_g, m->w.dproj_ln_b, nullptr, nullptr, nullptr,
                              nullptr, db.buf[0], s)) return 1; }
        return decoder_layers(m, B, T, db, 0, nullptr, mel, s);
    }
    if (m->use_tensor_core && m->w.dproj_w_h16 && m->dx2 == 256 && umma_dec256_supported(m->dx4, 5, 256, 2)) {
        { ProfRange r(ES_K_DEC_PROJ, s);
          if (launch_umma_dec256(2, B, T, m->dx4, m->dx2, features, nullptr, nullptr, m->w.dproj_w_h16,
                                 m->w.dproj_b, 1, m->w.dproj_ln_g, m->w.dproj_ln_b, nullptr, nullptr, nullptr,
                                 nullptr, db.buf[0], s)) return 1; }
        return decoder_layers(m, B, T, db, 0, nullptr, mel, s);
    }
    RowGemmParams p = base_params(B, T, T, m->dx4, m->dx2, features, m->dx4, m->w.dproj_w, db.buf[0], m->dx2);
    p.bias = m->w.dproj_b; p.act1 = ACT_TANH; p.ln_g = m->w.dproj_ln_g; p.ln_b = m->w.dproj_ln_b;
    { ProfRange r(ES_K_DEC_PROJ, s); if (launch_rowgemm(p, s)) return 1; }
    return decoder_layers(m, B, T, db, 0, nullptr, mel, s);
}

int es_decoder_forward_gathered(es_model_t* m, void* stream, int B, int N, int T,
                                const float* fused4, const int32_t* dur_cum, const int32_t* mel_len,
                                int zero_padded_frames, float* mel, void* workspace, size_t workspace_bytes) {
    ES_CHECK(m, "null model");
    ES_CHECK(B >= 1 && B <= 65535 && N >= 1 && T >= 1, "need 1 <= B <= 65535, N >= 1 and T >= 1 frames");
    ES_CHECK(fused4 && dur_cum && mel_len && mel, "null tensor");
    ES_CHECK(workspace && workspace_bytes >= decoder_workspace_bytes(m, B, N, T), "workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    Arena a(reinterpret_cast<void*>(align_up(reinterpret_cast<size_t>(workspace), 256)));
    DecBufs db = plan_decoder(m, a, B, N, T);
    // The projection is row-wise and the length regulator is a row gather: project once per PHONEME
    // (B*N rows), expand through the frame -> row map (es_gather.cu).  networks.py:228-258, :292
    const int R = B * N;
    { ProfRange r(ES_K_DEC_PROJ, s); if (project_rows(m, R, fused4, db.P, s)) return 1; }
    // Ragged schedule: with padded frames zeroed at the end (networks.py:424-427, B > 1) a tile that starts at or beyond
    // mel_len[b] + 2 L cannot reach a frame anyone reads (a layer looks 2 frames each way): the tcgen05 layer kernels
    // walk the compacted list of the other tiles, and the frames nobody computes are zero-filled in the mel -- both by
    // the index-map kernel, no extra launch.
    const bool ragged = zero_padded_frames && m->ragged_schedule && m->cfg.n_mel % 4 == 0 &&
                        (decoder_all_umma128(m) || decoder_all_umma256(m));
    const int tile_frames = m->dx2 == 128 ? 64 : 128, halo = (m->cfg.decoder_kernel_size / 2) * m->n_layers;
    { ProfRange r(ES_K_LENREG, s);
      if (launch_frame_source(dur_cum, mel_len, db.src, B, N, T, m->w.dproj_b, m->w.dproj_ln_g, m->w.dproj_ln_b,
                              m->dx2, db.P + (size_t)R * m->dx2, s, ragged ? db.tiles : nullptr, db.tile_count,
                              tile_frames, halo, mel, m->cfg.n_mel)) return 1; }
    int rc;
    if (m->gather_mode == ES_GATHER_FUSED && (decoder_all_umma128(m) || decoder_all_umma256(m))) {
        // the first block reads its input and skip rows straight from the table: [B,T,dx2] is never materialised
        rc = decoder_layers(m, B, T, db, -1, zero_padded_frames ? mel_len : nullptr, mel, s, ragged);
    } else {
        { ProfRange r(ES_K_LENREG, s); if (launch_gather_rows(db.P, db.src, db.buf[0], (long long)B * T, m->dx2, s)) return 1; }
        rc = decoder_layers(m, B, T, db, 0, zero_padded_frames ? mel_len : nullptr, mel, s, ragged);
    }
    return rc;
}

}  // extern "C"

namespace {

// every depthwise layer runs on the 128-channel tcgen05 kernel (which has the gathered-row variant)
bool decoder_all_umma128(const es_model* m) {
    if (!m->use_tensor_core || !umma_dec_supported(m->dx2, m->cfg.decoder_kernel_size, m->dx2)) return false;
    for (int l = 0; l < m->n_layers; ++l) if (!m->w.dec[l].pw_w_h16) return false;
    return true;
}

// every depthwise layer and the mel head run on the 256-channel tcgen05 kernel
bool decoder_all_umma256(const es_model* m) {
    if (!m->use_tensor_core || m->dx2 != 256 || !umma_dec256_supported(256, m->cfg.decoder_kernel_size, 256, 0)) return false;
    if (!m->w.mel_w_h16 || m->cfg.n_mel != 80) return false;
    for (int l = 0; l < m->n_layers; ++l) if (!m->w.dec[l].pw_w_h16) return false;
    return true;
}

// out[r] = LN(tanh(Linear(in[r]))) for `rows` independent rows (MelDecoder.proj, networks.py:292)
int project_rows(const es_model* m, int rows, const float* in, float* out, cudaStream_t s) {
    if (m->use_tensor_core && m->w.dproj_w_h16 && m->dx4 == 128 &&
        umma_dec_supported(m->dx2, m->cfg.decoder_kernel_size, m->dx2))
        return launch_umma_dec(2, 1, rows, m->dx2, in, nullptr, nullptr, m->w.dproj_w_h16,
                               m->w.dproj_b, 1, m->w.dproj_ln_g, m->w.dproj_ln_b, nullptr, nullptr, nullptr,
                               nullptr, out, s);
    if (m->use_tensor_core && m->w.dproj_w_h16 && m->dx2 == 256 && umma_dec256_supported(m->dx4, 5, 256, 2))
        return launch_umma_dec256(2, 1, rows, m->dx4, m->dx2, in, nullptr, nullptr, m->w.dproj_w_h16,
                                  m->w.dproj_b, 1, m->w.dproj_ln_g, m->w.dproj_ln_b, nullptr, nullptr, nullptr,
                                  nullptr, out, s);
    RowGemmParams p = base_params(1, rows, rows, m->dx4, m->dx2, in, m->dx4, m->w.dproj_w, out, m->dx2);
    p.bias = m->w.dproj_b; p.act1 = ACT_TANH; p.ln_g = m->w.dproj_ln_g; p.ln_b = m->w.dproj_ln_b;
    return launch_rowgemm(p, s);
}

// Decoder blocks + mel head (networks.py:293-302, :424-427).  db.buf[s_idx] holds `skip`; s_idx < 0: `skip` is
// virtual -- row db.src[b*T + t] of the table db.P -- and the first block gathers it (decoder_all_umma128 only).
int decoder_layers(const es_model* m, int B, int T, DecBufs& db, int s_idx, const int* zero_from,
                   float* mel, cudaStream_t s, bool ragged) {
    const int C = m->dx2;
    const int2* tl = ragged ? db.tiles : nullptr;             // ragged schedule (tcgen05 kernels only)
    const int* tc = ragged ? db.tile_count : nullptr;
    int layer = 0;
    for (int blk = 0; blk < m->cfg.n_blocks; ++blk) {
        int in_idx = s_idx;
        for (int l = 0; l < m->cfg.block_depth; ++l, ++layer) {
            int out_idx = 0;
            while (out_idx == s_idx || out_idx == in_idx) ++out_idx;
            const es_dec_layer_w_t& w = m->w.dec[layer];
            const bool last = (l == m->cfg.block_depth - 1);
            const bool gx = in_idx < 0, gs = last && s_idx < 0;
            if (gx || gs)
                ES_CHECK(db.P && db.src && w.pw_w_h16 && m->use_tensor_core &&
                             (umma_dec_supported(C, m->cfg.decoder_kernel_size, C) ||
                              (C == 256 && umma_dec256_supported(C, m->cfg.decoder_kernel_size, C, 0))),
                         "virtual skip needs a tcgen05 layer kernel");
            if (m->use_tensor_core && w.pw_w_h16 && umma_dec_supported(C, m->cfg.decoder_kernel_size, C)) {
                // 128-channel decoders: one fused tcgen05 kernel per layer; in the first block of the gathered entry the
                // input and / or skip rows come from the projection table through the frame -> row map
                ProfRange r(ES_K_DEC_LAYER, s);
                const float* xin = gx ? db.P : db.buf[in_idx];
                const float* skip = last ? (s_idx < 0 ? db.P : db.buf[s_idx]) : nullptr;
                int rc;
                if (gx || gs)
                    rc = launch_umma_dec_gathered(B, T, C, xin, w.dw_w, w.dw_b, w.pw_w_h16, w.pw_b, 1, w.ln_g, w.ln_b, skip,
                                                  last ? m->w.blk_ln_g[blk] : nullptr, last ? m->w.blk_ln_b[blk] : nullptr,
                                                  db.src, db.pad_id, gx, gs, db.buf[out_idx], s, tl, tc);
                else
                    rc = launch_umma_dec(0, B, T, C, xin, w.dw_w, w.dw_b, w.pw_w_h16, w.pw_b, 1, w.ln_g, w.ln_b, skip,
                                         last ? m->w.blk_ln_g[blk] : nullptr, last ? m->w.blk_ln_b[blk] : nullptr,
                                         nullptr, db.buf[out_idx], s, tl, tc);
                if (rc) return 1;
                in_idx = out_idx;
                continue;
            }
            if (m->use_tensor_core && w.pw_w_h16 && C == 256 && umma_dec256_supported(C, m->cfg.decoder_kernel_size, C, 0)) {
                ProfRange r(ES_K_DEC_LAYER, s);
                int rc;
                if (gx || gs) {
                    const float* xin = gx ? db.P : db.buf[in_idx];
                    const float* skip = last ? (s_idx < 0 ? db.P : db.buf[s_idx]) : nullptr;
                    rc = launch_umma_dec256_gathered(B, T, xin, w.dw_w, w.dw_b, w.pw_w_h16, w.pw_b, w.ln_g, w.ln_b, skip,
                                                     last ? m->w.blk_ln_g[blk] : nullptr, last ? m->w.blk_ln_b[blk] : nullptr,
                                                     db.src, gx, gs, db.buf[out_idx], s, tl, tc);
                } else {
                    rc = launch_umma_dec256(0, B, T, C, C, db.buf[in_idx], w.dw_w, w.dw_b, w.pw_w_h16,
                                            w.pw_b, 1, w.ln_g, w.ln_b, last ? db.buf[s_idx] : nullptr,
                                            last ? m->w.blk_ln_g[blk] : nullptr, last ? m->w.blk_ln_b[blk] : nullptr,
                                            nullptr, db.buf[out_idx], s, tl, tc);
                }
                if (rc) return 1;
                in_idx = out_idx;
                continue;
            }
            RowGemmParams p = base_params(B, T, T, C, C, db.buf[in_idx], C, w.pw_w, db.buf[out_idx], C);
            p.mode = ROW_DWCONV; p.dw_w = w.dw_w; p.dw_b = w.dw_b; p.dw_k = m->cfg.decoder_kernel_size;
            p.bias = w.pw_b; p.act1 = ACT_TANH; p.ln_g = w.ln_g; p.ln_b = w.ln_b;
            if (l == m->cfg.block_depth - 1) {       // skip = LN_blk(x + skip)   networks.py:299
                p.res2 = db.buf[s_idx]; p.ldr2 = C; p.ln2_g = m->w.blk_ln_g[blk]; p.ln2_b = m->w.blk_ln_b[blk];
            }
            { ProfRange r(ES_K_DEC_LAYER, s); if (launch_rowgemm(p, s)) return 1; }
            in_idx = out_idx;
        }
        s_idx = in_idx;
    }
    // mel = Linear(skip); padded frames zeroed                                      networks.py:302, :424-427
    if (m->use_tensor_core && m->w.mel_w_h16 && m->cfg.n_mel == 80 &&
        umma_dec_supported(C, m->cfg.decoder_kernel_size, m->cfg.n_mel)) {
        ProfRange r(ES_K_MEL, s);
        return launch_umma_dec(2, B, T, m->cfg.n_mel, db.buf[s_idx], nullptr, nullptr,
                               m->w.mel_w_h16, m->w.mel_b, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                               zero_from, mel, s, tl, tc);
    }
    if (m->use_tensor_core && m->w.mel_w_h16 && C == 256 && m->cfg.n_mel == 80 &&
        umma_dec256_supported(C, m->cfg.decoder_kernel_size, 80, 2)) {
        ProfRange r(ES_K_MEL, s);
        return launch_umma_dec256(2, B, T, C, m->cfg.n_mel, db.buf[s_idx], nullptr, nullptr,
                                  m->w.mel_w_h16, m->w.mel_b, 0, nullptr, nullptr, nullptr, nullptr, nullptr,
                                  zero_from, mel, s, tl, tc);
    }
    RowGemmParams p = base_params(B, T, T, C, m->cfg.n_mel, db.buf[s_idx], C, m->w.mel_w, mel, m->cfg.n_mel);
    p.bias = m->w.mel_b; p.zero_from = zero_from;
    ProfRange r(ES_K_MEL, s);
    return launch_rowgemm(p, s);
}

}  // namespace

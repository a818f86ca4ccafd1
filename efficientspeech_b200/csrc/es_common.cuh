// Shared device/host helpers for the EfficientSpeech sm_100a kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/es_b200.h"

namespace es {

constexpr float kLnEps = 1e-5f;   // nn.LayerNorm default eps (layers/networks.py:45-46)

// ---- error plumbing (host) -----------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launches;

#define ES_CHECK(cond, msg)                                                   \
    do {                                                                      \
        if (!(cond)) {                                                        \
            ::es::set_error(std::string(__func__) + ": " + (msg));            \
            return 1;                                                         \
        }                                                                     \
    } while (0)

#define ES_CUDA(expr)                                                         \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) {                                              \
            ::es::set_error(std::string(__func__) + ": " #expr " -> " +       \
                            cudaGetErrorString(_e));                          \
            return 1;                                                         \
        }                                                                     \
    } while (0)

#define ES_LAUNCH_OK()                                                        \
    do {                                                                      \
        ::es::g_launches.fetch_add(1, std::memory_order_relaxed);             \
        ES_CUDA(cudaGetLastError());                                          \
    } while (0)

// ---- per-device caches -------------------------------------------------------------------------------
// cudaFuncSetAttribute, the SM count and device allocations belong to ONE device; a process may drive several
// (the Python API accepts any device), so everything cached across calls is keyed by cudaGetDevice().
constexpr int kMaxDevices = 64;
template <typename T>
struct PerDeviceSlot {
    T v[kMaxDevices] = {};
    T& get() {
        int d = 0;
        cudaGetDevice(&d);
        return v[d & (kMaxDevices - 1)];
    }
};

// ---- launch with programmatic stream serialization (PDL); the kernel must call griddepcontrol.wait --
template <typename Param>
inline cudaError_t launch_pdl(void (*kernel)(Param), dim3 grid, int block, size_t smem, cudaStream_t s, const Param& p) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

// ---- optional per-launch event timing (es_profile_*) --------------------------------------
void prof_begin_range(int kind, cudaStream_t s);
void prof_end_range(cudaStream_t s);
struct ProfRange {
    cudaStream_t s;
    ProfRange(int kind, cudaStream_t st) : s(st) { prof_begin_range(kind, st); }
    ~ProfRange() { prof_end_range(s); }
};

// ---- activations -----------------------------------------------------------------------
enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_TANH = 3 };

// Out-of-line transcendental activations: epilogue loops stay fully unrolled (values in registers)
// without replicating the erf / tanh expansions per element (instruction-cache footprint).
static __device__ __noinline__ float gelu_erf_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }  // blocks.py:19
static __device__ __noinline__ float tanh_f(float v) { return tanhf(v); }

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == ACT_RELU) return fmaxf(v, 0.f);
    if (act == ACT_GELU) return gelu_erf_f(v);
    if (act == ACT_TANH) return tanh_f(v);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- generic fused row-GEMM (es_rowgemm.cu) ----------------------------------------------
enum RowMode : int { ROW_PLAIN = 0, ROW_DWCONV = 2 };

struct RowGemmParams {
    // geometry: output rows are (b, t), t in [0, n_out); input rows (b, t_in), t_in in [0, n_in)
    int B, n_in, n_out;
    int K, Nout, ldw;          // ldw: row stride of W (Nout padded to a multiple of 32)
    int taps, stride, pad;     // PLAIN: t_in = t*stride + tau - pad, zero outside [0, n_in)
    int mode;
    const float* A;            // [B*n_in][lda]
    int lda;
    // DWCONV: A'[t][c] = dw_b[c] + sum_tau dw_w[tau][c] * A[t + tau - dw_k/2][c]
    const float* dw_w;
    const float* dw_b;
    int dw_k;
    // weights
    const float* W;            // [taps][K][ldw]
    const float* bias;         // [ldw] or null
    const float* tap_bias;     // [taps][ldw] or null: added where tap tau reads inside the sequence
    int act1;
    // optional scalar head on the post-act1 values: dot_out[row] = (relu?)(sum_c v[c] dot_w[c] + dot_b[0])
    const float* dot_w;
    const float* dot_b;
    float* dot_out;
    int dot_relu;
    const float* res1;         // added before LN1, [B*n_out][ldr1]
    int ldr1;
    const float* ln_g;         // LayerNorm over Nout (requires one column tile) or null
    const float* ln_b;
    int act2;
    const float* res2;         // out = LN2(out + res2)
    int ldr2;
    const float* ln2_g;
    const float* ln2_b;
    // Fuse epilogue (es_umma_wide.cu only): out[t] += sum over tau = t mod 2, tau < fuse_k, 0 <= (t-tau)/2 < fuse_n1
    // of fuse_u[(b*fuse_n1 + (t-tau)/2) * fuse_ld + tau*Nout + col]         (ConvTranspose1d stride 2, networks.py:199-206)
    const float* fuse_u;
    int fuse_k, fuse_n1, fuse_ld;
    const uint8_t* row_mask;   // [B][n_out], 1 -> zero the row
    const int* zero_from;      // [B], rows t >= zero_from[b] are zeroed
    float* Y;                  // [B*n_out][ldy] or null
    int ldy;
};

int launch_rowgemm(const RowGemmParams& p, cudaStream_t stream);
int launch_rowgemm_narrow(const RowGemmParams& p, cudaStream_t stream);   // -1: not applicable
int launch_rowgemm_narrow_batch(const RowGemmParams* ps, int count, cudaStream_t stream);   // -1: not applicable

}  // namespace es

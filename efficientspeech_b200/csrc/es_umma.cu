// tcgen05 building block self-test: C[M,N] = A[M,K] . B[N,K]^T with split-fp16 operands.
// Validates, in isolation, everything the fused decoder kernel relies on: the canonical
// no-swizzle K-major shared-memory layout, the UMMA shared-memory and instruction
// descriptors, TMEM allocation, tcgen05.mma issue/commit on an mbarrier, and the
// tcgen05.ld 32x32b epilogue mapping (TMEM lane = tile row).
#include <stdlib.h>

#include "es_common.cuh"
#include "es_umma.cuh"

namespace es {
namespace {

using namespace umma;

__global__ void __launch_bounds__(128)
umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ C,
                     int N, int K, int variant, int* __restrict__ err) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_s;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * 128;
    const uint32_t a_bytes = 128u * K * 2u, b_bytes = (uint32_t)N * K * 2u;
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + a_bytes;
    uint8_t* b_hi = a_lo + a_bytes;
    uint8_t* b_lo = b_hi + b_bytes;
    const uint32_t ncols = N <= 32 ? 32u : N <= 64 ? 64u : N <= 128 ? 128u : 256u;
    const uint32_t bar = smem_u32(&bar_s);

    if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), ncols);
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }

    const int kchunks = K >> 3;
    for (int idx = tid; idx < 128 * kchunks; idx += 128) {
        const int row = idx & 127, kc = idx >> 7;
        const float4* src = reinterpret_cast<const float4*>(A + (size_t)(m0 + row) * K + kc * 8);
        const float4 p = __ldg(src), q = __ldg(src + 1);
        const float v[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
        uint4 hi, lo;
        split8(v, hi, lo);
        const uint32_t off = canon_off(row, kc, 128);
        *reinterpret_cast<uint4*>(a_hi + off) = hi;
        *reinterpret_cast<uint4*>(a_lo + off) = lo;
    }
    for (int idx = tid; idx < N * kchunks; idx += 128) {
        const int row = idx % N, kc = idx / N;
        const float4* src = reinterpret_cast<const float4*>(Bm + (size_t)row * K + kc * 8);
        const float4 p = __ldg(src), q = __ldg(src + 1);
        const float v[8] = {p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w};
        uint4 hi, lo;
        split8(v, hi, lo);
        const uint32_t off = canon_off(row, kc, N);
        *reinterpret_cast<uint4*>(b_hi + off) = hi;
        *reinterpret_cast<uint4*>(b_lo + off) = lo;
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {
        const uint32_t idesc = make_idesc_f16(128, N);
        uint32_t lbo_a = 128u * 16u, sbo_a = 128u, lbo_b = (uint32_t)N * 16u, sbo_b = 128u;
        if (variant & 1) { uint32_t t = lbo_a; lbo_a = sbo_a; sbo_a = t; t = lbo_b; lbo_b = sbo_b; sbo_b = t; }
        for (int s = 0; s < (K >> 4); ++s) {
            const uint32_t ao = (uint32_t)(2 * s) * 128u * 16u, bo = (uint32_t)(2 * s) * (uint32_t)N * 16u;
            const uint64_t dah = make_smem_desc(smem_u32(a_hi) + ao, lbo_a, sbo_a);
            const uint64_t dal = make_smem_desc(smem_u32(a_lo) + ao, lbo_a, sbo_a);
            const uint64_t dbh = make_smem_desc(smem_u32(b_hi) + bo, lbo_b, sbo_b);
            const uint64_t dbl = make_smem_desc(smem_u32(b_lo) + bo, lbo_b, sbo_b);
            mma_f16_ss(tmem, dah, dbh, idesc, s > 0 ? 1u : 0u);
            mma_f16_ss(tmem, dah, dbl, idesc, 1u);
            mma_f16_ss(tmem, dal, dbh, idesc, 1u);
        }
        mma_commit(bar);
    }
    const bool ok = mbar_wait(bar, 0);
    if (!ok && lane == 0) atomicExch(err, 1);
    tc_fence_after_sync();

    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        float* dst = C + (size_t)(m0 + row) * N + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                              __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

}  // namespace
}  // namespace es

extern "C" int es_selftest_umma_gemm(void* stream, int M, int N, int K, const float* A, const float* Bm, float* C) {
    using namespace es;
    ES_CHECK(A && Bm && C, "null tensor");
    ES_CHECK(M > 0 && M % 128 == 0, "M must be a multiple of 128");
    ES_CHECK(N % 16 == 0 && N >= 32 && N <= 256, "N must be a multiple of 16 in [32, 256]");
    ES_CHECK(K % 16 == 0 && K >= 16, "K must be a multiple of 16");
    const size_t smem = (size_t)(128 + N) * K * 4;
    ES_CHECK(smem <= 200 * 1024, "operands do not fit in shared memory");
    const char* v = getenv("ES_UMMA_VARIANT");
    const int variant = v ? atoi(v) : 0;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    static int* d_err = nullptr;
    if (!d_err) ES_CUDA(cudaMalloc(&d_err, sizeof(int)));
    ES_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), s));
    ES_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    umma_selftest_kernel<<<M / 128, 128, smem, s>>>(A, Bm, C, N, K, variant, d_err);
    ES_LAUNCH_OK();
    int h_err = 0;
    ES_CUDA(cudaMemcpyAsync(&h_err, d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    ES_CUDA(cudaStreamSynchronize(s));
    ES_CHECK(h_err == 0, "tcgen05.commit never arrived on the mbarrier (timeout)");
    return 0;
}

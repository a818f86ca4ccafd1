// tcgen05 / TMEM / mbarrier / bulk-copy primitives (sm_100a inline PTX) shared by the tensor-core
// kernels.  Operand convention used throughout this repo:
//
//   * operands are fp16 "split" pairs: x = hi + lo with hi = fp16(x), lo = fp16(x - hi), so
//     hi*hi' + hi*lo' + lo*hi' (3 tcgen05.mma, fp32 accumulate in TMEM) reproduces the fp32
//     product to ~2^-21 relative -- single-pass TF32/BF16 misses the 1e-3 mel bar;
//   * shared-memory operand tiles use the UMMA canonical K-major NO-SWIZZLE ("interleave")
//     layout: 8x8 core matrices of 128 contiguous bytes (8 rows x 16 B),
//         addr(row, k) = (k/8)*LBO + (row/8)*SBO + (row%8)*16 + (k%8)*2          [bytes, fp16]
//     with SBO = 128 (8-row groups contiguous) and LBO = rows*16 (one [rows x 8] K-panel after
//     the other).  This layout is trivially produced by CUDA cores (the decoder's A operand is
//     COMPUTED -- depthwise conv -- not loaded), and 32 lanes writing 32 consecutive rows of one
//     K-panel hit 512 contiguous bytes: bank-conflict free.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace es {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// named barrier for a sub-group of warps (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// suspend-time hint: a waiting thread may sleep in hardware until the phase completes (or this many
// ns pass) instead of re-issuing the test -- the spinning warps otherwise compete for issue slots
#ifndef ES_MBAR_SUSPEND_NS
#define ES_MBAR_SUSPEND_NS 20000
#endif
constexpr uint32_t kMbarSuspendNs = ES_MBAR_SUSPEND_NS;      // 0: plain try_wait (implementation-defined short limit)
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    if (kMbarSuspendNs != 0) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity), "r"(kMbarSuspendNs) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
    return ok != 0;
}
// Bounded wait: a wrong descriptor must not hang the GPU.  Returns false on timeout (~seconds).
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t i = 0; i < (1u << 26); ++i)
        if (mbar_try_wait(bar, parity)) return true;
    return false;
}

// One lane of a fully converged warp (elect.sync).  Code guarded by it stays warp-uniform for the
// compiler, so UMMA descriptors / barrier addresses are computed in UNIFORM registers instead of
// being moved there (R2UR) one by one in front of every tcgen05.mma.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor in the stream is still draining; everything before pdl_wait() (TMEM allocation,
// barrier init, parameter / weight staging) overlaps the predecessor's tail, everything after sees
// the predecessor's global writes.  pdl_launch_dependents() lets the successor start its own prologue.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- proxies / fences -------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {      // generic st.shared -> visible to UMMA/TMA
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- bulk async copy (TMA 1-D): global -> shared, completes on an mbarrier ----------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// shared -> global bulk store (TMA 1-D); completion is tracked per thread in bulk groups
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }   // sources may be overwritten
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }             // writes are complete

// ---- Ampere-style async copies (16 B, zero-fill when src_bytes == 0) ------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMEM -----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // whole warp, ncols pow2 >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of warp w gets TMEM lane 32*(w%4)+i.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
// narrower forms of the same shape: 8 / 16 consecutive fp32 columns of the thread's TMEM lane
__device__ __forceinline__ void tmem_ld32x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
// 16 lanes x 256 bit, repeated 16x along columns: a warp reads 16 TMEM lanes (rows) x 128 fp32
// columns.  Thread t gets, for column group j (8 columns): r[4j+0..1] = row t/4, cols 8j+2(t%4)+{0,1};
// r[4j+2..3] = row t/4+8, same columns (cute Copy_Traits<SM100_TMEM_LOAD_16dp256b*>), i.e. the
// mma.sync C-fragment layout: 4 threads share a row, and one warp-wide 8-byte access touches
// 8 rows x 32 contiguous bytes (full 32-byte sectors).
__device__ __forceinline__ void tmem_ld_16x256b_x16(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
          "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr) : "memory");
}
// store counterpart of tmem_ld_16x256b_x16 (same thread <-> element mapping)
__device__ __forceinline__ void tmem_st_16x256b_x16(uint32_t taddr, const uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x16.b32 [%64], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63};"
        :: "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
           "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
           "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
           "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]),
           "r"(r[32]), "r"(r[33]), "r"(r[34]), "r"(r[35]), "r"(r[36]), "r"(r[37]), "r"(r[38]), "r"(r[39]),
           "r"(r[40]), "r"(r[41]), "r"(r[42]), "r"(r[43]), "r"(r[44]), "r"(r[45]), "r"(r[46]), "r"(r[47]),
           "r"(r[48]), "r"(r[49]), "r"(r[50]), "r"(r[51]), "r"(r[52]), "r"(r[53]), "r"(r[54]), "r"(r[55]),
           "r"(r[56]), "r"(r[57]), "r"(r[58]), "r"(r[59]), "r"(r[60]), "r"(r[61]), "r"(r[62]), "r"(r[63]),
           "r"(taddr) : "memory");
}
// 64-column variants (32 registers): reg 4j+2i+b = row (lane/4 + 8i), column 8j + 2(lane%4) + b, j < 8
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x8(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}
// 32-column variants (16 registers): reg 4j+2i+b = row (lane/4 + 8i), column 8j + 2(lane%4) + b, j < 4
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
// per-warpgroup register re-allocation (all 4 warps of an aligned warpgroup execute it; N multiple of 8 in [24, 256])
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout_type=0
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// Instruction descriptor, kind::f16: fp16 x fp16 -> fp32, both operands K-major
// (cute::UMMA::InstrDescriptor: c_format [4,6)=1 (F32), a/b_format = 0 (F16), n_dim [17,23) = N>>3,
//  m_dim [24,29) = M>>4)
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T, issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- split-fp16 helpers -----------------------------------------------------------------------------
// 8 consecutive-k fp32 values of one row -> two 16-byte core-matrix rows (hi, lo).
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
    __half2 h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 back = __half22float2(hh);
        h[i] = hh;
        l[i] = __floats2half2_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);
    }
    hi = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                    *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
    lo = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]),
                    *reinterpret_cast<uint32_t*>(&l[2]), *reinterpret_cast<uint32_t*>(&l[3]));
}

// byte offset of (row, 8-wide k chunk) inside a canonical no-swizzle tile of `rows` rows
__device__ __forceinline__ uint32_t canon_off(int row, int kchunk, int rows) {
    return (uint32_t)kchunk * (uint32_t)(rows * 16) + (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u;
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2) ------------------------------------------------------
// Two fp32 lanes in one 64-bit register pair, one issue slot per instruction: the epilogues and the
// operand producers of the tensor-core kernels are issue-bound, not FLOP-bound, so halving the FMA-pipe
// instruction count is what matters.  Each lane is an ordinary IEEE fma.rn / add.rn (bit-identical to
// the scalar form).  pk2/up2 are register-allocation no-ops when the halves already sit in an aligned pair.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ f32x2 pk2u(uint32_t lo, uint32_t hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ float2 up2(f32x2 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---- shared epilogue / prologue math ------------------------------------------------------------------
constexpr float kLnEpsU = 1e-5f;
// 2 consecutive channels -> split fp16 pair: hi = fp16(v), lo = fp16(v - hi)
__device__ __forceinline__ void split2(f32x2 v, uint32_t& hi, uint32_t& lo) {
    const float2 a = up2(v);
    const __half2 h = __floats2half2_rn(a.x, a.y);
    const float2 b = __half22float2(h);
    const float2 d = up2(sub2(v, pk2(b.x, b.y)));
    const __half2 l = __floats2half2_rn(d.x, d.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
// 4 consecutive channels -> split fp16: hi = fp16(v), lo = fp16(v - hi)
__device__ __forceinline__ void split4(float4 v, uint2& hi, uint2& lo) {
    split2(pk2(v.x, v.y), hi.x, lo.x);
    split2(pk2(v.z, v.w), hi.y, lo.y);
}
__device__ __forceinline__ void split4(ulonglong2 v, uint2& hi, uint2& lo) {
    split2(v.x, hi.x, lo.x);
    split2(v.y, hi.y, lo.y);
}

// LayerNorm in the 16x256b fragment layout: v[4j+2*row+b] = column 8j+2*t4+b of this thread's row
// `row`; the 4 threads of a quad (t4 = 0..3) hold one row.  Single pass (sum and sum of squares in
// two independent chains each -- the inputs are tanh / LayerNorm outputs of O(1) magnitude, so
// E[x^2]-m^2 loses nothing at fp32), 2 xor-shuffles per statistic, then y = ((v*r - m*r) * g + b):
// 4 FMA-pipe instructions per element.
__device__ __forceinline__ void quad_stats(float s, float q, float inv_n, float& r, float& nm) {
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    q += __shfl_xor_sync(0xffffffffu, q, 2);
    const float m = s * inv_n;
    r = rsqrtf(fmaf(q, inv_n, -m * m) + kLnEpsU);
    nm = -m * r;
}

// both rows of the thread (shares the gamma/beta loads)
__device__ __forceinline__ void fragment_layernorm2(float (&v)[64], const float* __restrict__ g,
                                                    const float* __restrict__ be, int t4, float inv_n) {
    float s0a = 0.f, s0b = 0.f, q0a = 0.f, q0b = 0.f, s1a = 0.f, s1b = 0.f, q1a = 0.f, q1b = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        s0a += v[4 * j]; s0b += v[4 * j + 1]; q0a = fmaf(v[4 * j], v[4 * j], q0a); q0b = fmaf(v[4 * j + 1], v[4 * j + 1], q0b);
        s1a += v[4 * j + 2]; s1b += v[4 * j + 3]; q1a = fmaf(v[4 * j + 2], v[4 * j + 2], q1a); q1b = fmaf(v[4 * j + 3], v[4 * j + 3], q1b);
    }
    float r0, n0, r1, n1;
    quad_stats(s0a + s0b, q0a + q0b, inv_n, r0, n0);
    quad_stats(s1a + s1b, q1a + q1b, inv_n, r1, n1);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float2 gg = *reinterpret_cast<const float2*>(g + 8 * j + 2 * t4);
        const float2 bb = *reinterpret_cast<const float2*>(be + 8 * j + 2 * t4);
        v[4 * j] = fmaf(fmaf(v[4 * j], r0, n0), gg.x, bb.x);
        v[4 * j + 1] = fmaf(fmaf(v[4 * j + 1], r0, n0), gg.y, bb.y);
        v[4 * j + 2] = fmaf(fmaf(v[4 * j + 2], r1, n1), gg.x, bb.x);
        v[4 * j + 3] = fmaf(fmaf(v[4 * j + 3], r1, n1), gg.y, bb.y);
    }
}

// one row (ROW = 0 | 1) of the thread
template <int ROW>
__device__ __forceinline__ void fragment_layernorm_row(float (&v)[64], const float* __restrict__ g,
                                                       const float* __restrict__ be, int t4, float inv_n) {
    float sa = 0.f, sb = 0.f, qa = 0.f, qb = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        sa += v[4 * j + 2 * ROW]; sb += v[4 * j + 2 * ROW + 1];
        qa = fmaf(v[4 * j + 2 * ROW], v[4 * j + 2 * ROW], qa);
        qb = fmaf(v[4 * j + 2 * ROW + 1], v[4 * j + 2 * ROW + 1], qb);
    }
    float r, nm;
    quad_stats(sa + sb, qa + qb, inv_n, r, nm);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float2 gg = *reinterpret_cast<const float2*>(g + 8 * j + 2 * t4);
        const float2 bb = *reinterpret_cast<const float2*>(be + 8 * j + 2 * t4);
        v[4 * j + 2 * ROW] = fmaf(fmaf(v[4 * j + 2 * ROW], r, nm), gg.x, bb.x);
        v[4 * j + 2 * ROW + 1] = fmaf(fmaf(v[4 * j + 2 * ROW + 1], r, nm), gg.y, bb.y);
    }
}

// ---- packed-pair versions of the fragment LayerNorm: v[2j + row] holds columns 8j+2*t4, +1 of this thread's
// row `row` (the 16x256b fragment layout, adjacent columns paired) ---------------------------------------
__device__ __forceinline__ void fragment_layernorm2_p(f32x2 (&v)[32], const float* __restrict__ g,
                                                      const float* __restrict__ be, int t4, float inv_n) {
    f32x2 s0 = 0ull, q0 = 0ull, s1 = 0ull, q1 = 0ull;          // (+0.f, +0.f)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        s0 = add2(s0, v[2 * j]); q0 = fma2(v[2 * j], v[2 * j], q0);
        s1 = add2(s1, v[2 * j + 1]); q1 = fma2(v[2 * j + 1], v[2 * j + 1], q1);
    }
    const float2 a0 = up2(s0), b0 = up2(q0), a1 = up2(s1), b1 = up2(q1);
    float r0, n0, r1, n1;
    quad_stats(a0.x + a0.y, b0.x + b0.y, inv_n, r0, n0);
    quad_stats(a1.x + a1.y, b1.x + b1.y, inv_n, r1, n1);
    const f32x2 R0 = pk2(r0, r0), N0 = pk2(n0, n0), R1 = pk2(r1, r1), N1 = pk2(n1, n1);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const f32x2 gg = *reinterpret_cast<const f32x2*>(g + 8 * j + 2 * t4);
        const f32x2 bb = *reinterpret_cast<const f32x2*>(be + 8 * j + 2 * t4);
        v[2 * j] = fma2(fma2(v[2 * j], R0, N0), gg, bb);
        v[2 * j + 1] = fma2(fma2(v[2 * j + 1], R1, N1), gg, bb);
    }
}
template <int ROW>
__device__ __forceinline__ void fragment_layernorm_row_p(f32x2 (&v)[32], const float* __restrict__ g,
                                                         const float* __restrict__ be, int t4, float inv_n) {
    f32x2 s = 0ull, q = 0ull, s2 = 0ull, q2 = 0ull;            // two chains per statistic (one row only: less ILP)
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
        s = add2(s, v[2 * j + ROW]); q = fma2(v[2 * j + ROW], v[2 * j + ROW], q);
        s2 = add2(s2, v[2 * j + 2 + ROW]); q2 = fma2(v[2 * j + 2 + ROW], v[2 * j + 2 + ROW], q2);
    }
    const float2 a = up2(add2(s, s2)), b = up2(add2(q, q2));
    float r, nm;
    quad_stats(a.x + a.y, b.x + b.y, inv_n, r, nm);
    const f32x2 R = pk2(r, r), NM = pk2(nm, nm);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const f32x2 gg = *reinterpret_cast<const f32x2*>(g + 8 * j + 2 * t4);
        const f32x2 bb = *reinterpret_cast<const f32x2*>(be + 8 * j + 2 * t4);
        v[2 * j + ROW] = fma2(fma2(v[2 * j + ROW], R, NM), gg, bb);
    }
}

}  // namespace umma
}  // namespace es

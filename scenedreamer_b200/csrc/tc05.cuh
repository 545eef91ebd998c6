// Thin inline-PTX layer over the Blackwell (sm_100a) tensor-core path used by the fused render
// kernel: tcgen05.mma (kind::f16, operands in shared memory, fp32 accumulators in TMEM), TMEM
// allocation / loads, mbarriers, bulk async copies (TMA 1-D) and the proxy fences between them.
//
// Operand layout (both A and B are K-major, SWIZZLE_NONE "interleaved" canonical layout):
//   the matrix is cut into 8-row x 16-byte core matrices, each stored as 128 contiguous bytes;
//   core matrices that are neighbours along M/N are SBO = 128 B apart, neighbours along K are
//   LBO = rows * 16 B apart.  Element (r, k) of a [rows, K] 16-bit matrix therefore lives at
//       (k / 8) * rows * 16  +  (r / 8) * 128  +  (r % 8) * 16  +  (k % 8) * 2      bytes,
//   i.e. for a fixed 8-wide k-chunk all rows form one contiguous slab of rows * 16 bytes.  A
//   thread that owns row r writes its 8 consecutive k-values with ONE 16-byte st.shared and a
//   warp (32 consecutive rows) writes 512 contiguous bytes -> conflict-free epilogue stores, and
//   a weight K-chunk is one contiguous range that a single cp.async.bulk can fetch.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc05 {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- descriptors -----------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor (SWIZZLE_NONE, version 1 = Blackwell):
//   [0,14) start address >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version, [61,64) layout.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// 32-bit instruction descriptor for kind::f16 with fp32 accumulation, both operands K-major:
//   [4,6) D format (1 = f32), [7,10) A format, [10,13) B format (0 = f16, 1 = bf16),
//   [15] A major (0 = K), [16] B major (0 = K), [17,23) N >> 3, [24,29) M >> 4.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t M, uint32_t N, bool bf16) {
    return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- MMA issue / completion --------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Make the mbarrier track completion of all MMAs issued so far by this thread (implies
// tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void fence_before_thread_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_thread_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM --------------------------------------------------------------------------------------
// One full warp allocates `cols` (power of two >= 32) columns; the base address lands in *slot.
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets row (lane base + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// For waits that are NOT on the critical path (producers running ahead): back off with nanosleep so
// the spinning warps do not steal issue slots from the warps doing the epilogue math on the same SMSP.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity, uint32_t ns = 256) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

// ---- 1-D bulk async copy global -> shared (TMA engine), completion on an mbarrier -------------
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- 16-bit packing ----------------------------------------------------------------------------
template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
    if constexpr (BF16) {
        __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&v);
    } else {
        __half2 v = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t *>(&v);
    }
}
template <bool BF16>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
    if constexpr (BF16) {
        return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&u));
    } else {
        return __half22float2(*reinterpret_cast<__half2 *>(&u));
    }
}

// byte offset of the 16-byte chunk holding elements (r, 8*kc .. 8*kc+7) of a [rows, K] operand
__host__ __device__ __forceinline__ constexpr uint32_t chunk_off(uint32_t rows, uint32_t r, uint32_t kc) {
    return kc * rows * 16u + (r >> 3) * 128u + (r & 7u) * 16u;
}

}  // namespace tc05

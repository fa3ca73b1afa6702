// rs_tc.cuh -- tcgen05 (5th-gen tensor core) building blocks for the PPO2 MLP GEMMs, sm_100a only.
//
// One CTA of 128 threads owns a 128-sample tile.  Operands are fp32 tiles in shared memory consumed as
// kind::tf32 (the tensor core reads the top 19 bits), accumulators live in TMEM (128 lanes x N columns, fp32) and are read
// back with tcgen05.ld so that thread t gets row t -- the same "one thread = one sample row" ownership the FP32 kernels
// of rs_learn.cuh use for their epilogues.
//
// Shared-memory operand layout ("core-matrix" layout, no swizzle): a tile of R rows x C columns (C % 4 == 0, R % 8 == 0)
// is stored as 8-row x 16-byte core matrices of 128 contiguous bytes,
//      byte_offset(r, c) = (r / 8) * (C / 4) * 128 + (c / 4) * 128 + (r % 8) * 16 + (c % 4) * 4 .
// The SAME bytes serve two descriptor views:
//   * K-major  (rows = M or N index, columns = K):  LBO = 128 (next 16-byte K chunk), SBO = (C/4)*128 (next 8 rows);
//     one MMA (K = 8 tf32) consumes two K chunks -> the start address advances by 256 bytes per K step;
//   * MN-major (columns = M or N index, rows = K):  SBO = 128 (next 4 MN elements), LBO = (C/4)*128 (next 8 K);
//     one MMA consumes 8 rows -> the start address advances by (C/4)*128 bytes per K step.
// This is what lets  dW = in^T * dz  read the forward activations in place.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rstc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int tile_off(int r, int c, int C) {      // float index inside a core-matrix tile with C columns
    return (r >> 3) * (C >> 2) * 32 + (c >> 2) * 32 + (r & 7) * 4 + (c & 3);
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type=0 (no swizzle) [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// 32-bit instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate
__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {     // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_slot)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {         // same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A * B, issued by ONE thread.  nk K-steps of 8; the operand start addresses advance by a_step / b_step bytes.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_step,
                                          uint32_t b_addr, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_step, uint32_t idesc,
                                          int nk, uint64_t* bar) {
    for (int k = 0; k < nk; k++) {
        const uint64_t da = make_desc(a_addr + k * a_step, a_lbo, a_sbo), db = make_desc(b_addr + k * b_step, b_lbo, b_sbo);
        const uint32_t acc = k > 0 ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// same, with the choice of overwriting or accumulating into D on the first K-step (K-split products)
__device__ __forceinline__ void umma_tf32_acc(uint32_t tmem_d, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_step,
                                              uint32_t b_addr, uint32_t b_lbo, uint32_t b_sbo, uint32_t b_step, uint32_t idesc,
                                              int nk, bool accumulate, uint64_t* bar) {
    for (int k = 0; k < nk; k++) {
        const uint64_t da = make_desc(a_addr + k * a_step, a_lbo, a_sbo), db = make_desc(b_addr + k * b_step, b_lbo, b_sbo);
        const uint32_t acc = (k > 0 || accumulate) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// thread t of the CTA reads accumulator row t % 128 (warp w owns TMEM lanes 32 (w % 4) .. +31), columns col0 .. col0+31
__device__ __forceinline__ void tmem_ld32(uint32_t tmem_base, int col0, float* out) {
    const uint32_t taddr = tmem_base + ((((threadIdx.x >> 5) & 3u) * 32u) << 16) + (uint32_t)col0;       // warp w may touch lanes 32 (w % 4) .. +31
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) out[i] = __uint_as_float(r[i]);
}

// ---- descriptor / layout self-test: D[128 x 64] = op(A) * op(B) with every descriptor field supplied by the caller ----
// prm = { a_rows, a_cols, b_rows, b_cols, a_mn, b_mn, a_lbo, a_sbo, a_step, b_lbo, b_sbo, b_step, nk }
struct SelfTestParams { int v[13]; };
__global__ void __launch_bounds__(128) k_tc_selftest(SelfTestParams P, const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    extern __shared__ __align__(16) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    float* As = sm;
    float* Bs = sm + 128 * 128;
    const int t = threadIdx.x;
    const int ar = P.v[0], ac = P.v[1], br = P.v[2], bc = P.v[3];
    for (int i = t; i < ar * ac; i += 128) { int r = i / ac, c = i % ac; As[tile_off(r, c, ac)] = A[i]; }
    for (int i = t; i < br * bc; i += 128) { int r = i / bc, c = i % bc; Bs[tile_off(r, c, bc)] = B[i]; }
    if (t == 0) mbar_init(&bar, 1);
    if (t < 32) tmem_alloc(&tmem_slot, 64);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_slot;
    if (t == 0)
        umma_tf32(tm, smem_u32(As), P.v[6], P.v[7], P.v[8], smem_u32(Bs), P.v[9], P.v[10], P.v[11], make_idesc(128, 64, P.v[4], P.v[5]), P.v[12], &bar);
    mbar_wait(&bar, 0);
    tc_fence_after();
    float o[32];
    for (int c0 = 0; c0 < 64; c0 += 32) {
        tmem_ld32(tm, c0, o);
        for (int c = 0; c < 32; c++) D[t * 64 + c0 + c] = o[c];
    }
    tc_fence_before();
    __syncthreads();
    if (t < 32) tmem_dealloc(tm, 64);
}

}  // namespace rstc

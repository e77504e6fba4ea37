// swapnet_b200 — shared device/host helpers (sm_100a only).
//
// Thin inline-PTX wrappers for the Blackwell primitives the kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld),
// plus the error plumbing of the C-ABI (include/swapnet_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

// ----------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns 0 or a negative code and
// leaves a message readable through sn_last_error().
// ----------------------------------------------------------------------------
void sn_set_error(const char* fmt, ...);
#define SN_OK 0
#define SN_ERR_INVALID (-1)
#define SN_ERR_CUDA (-2)
#define SN_ERR_UNSUPPORTED (-3)
#ifndef SN_FMT_BF16
#define SN_FMT_BF16 0
#define SN_FMT_F16 1
#endif

#define SN_CHECK_CUDA(expr)                                                        \
  do {                                                                             \
    cudaError_t _e = (expr);                                                       \
    if (_e != cudaSuccess) {                                                       \
      sn_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,           \
                   cudaGetErrorString(_e));                                        \
      return SN_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

#define SN_REQUIRE(cond, ...)                                                      \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      sn_set_error(__VA_ARGS__);                                                   \
      return SN_ERR_INVALID;                                                       \
    }                                                                              \
  } while (0)

// ----------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a wrong descriptor must turn into a trap (launch failure), not
// a hung GPU box.  ~2^31 polls is minutes; real waits are microseconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) {
      printf("swapnet_b200: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x,
             blockIdx.y, blockIdx.z, threadIdx.x);
      __trap();
    }
  }
}

// ---- TMA ----------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* desc, uint64_t* bar, void* dst, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const void* desc, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(const void* desc, uint64_t* bar, void* dst, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---- tcgen05 ------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// warp-collective: allocate `ncols` TMEM columns, base address written to *dst (smem)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols)
               : "memory");
}
// single-thread issue: D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// single-thread issue: arrive on mbarrier when all previously issued MMAs retire
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// warp-collective: 32 lanes x 16 consecutive fp32 columns -> 16 regs per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- UMMA descriptors -----------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 |
//   [49,52) base_offset=0 | [61,64) layout (2 = SWIZZLE_128B)
//   layout codes (cute::UMMA::LayoutType): 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes, uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7u) << 61;
  return d;
}
// swizzle layout code of an operand whose rows are `chunk` 16-bit elements (64 / 32 / 16) wide
__host__ __device__ __forceinline__ uint32_t umma_layout_of_chunk(int chunk) {
  return chunk >= 64 ? 2u : (chunk == 32 ? 4u : 6u);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16, bf16 x bf16 -> f32.
//   c_format[4,6)=1(F32) a_format[7,10)=1(BF16) b_format[10,13)=1(BF16)
//   a_major bit15, b_major bit16 (0 = K-major, 1 = MN-major), n>>3 at [17,23), m>>4 at [24,29)
//   a_format / b_format: 0 = F16, 1 = BF16.  (Mixing the two in one instruction raises
//   "illegal instruction" on B200 — measured — so the host enforces a_fmt == b_fmt.)
__host__ __device__ __forceinline__ uint32_t umma_idesc_16(uint32_t m, uint32_t n, uint32_t a_fmt,
                                                           uint32_t b_fmt, uint32_t a_mn_major,
                                                           uint32_t b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (a_fmt == SN_FMT_F16 ? 0u : 1u) << 7;
  d |= (b_fmt == SN_FMT_F16 ? 0u : 1u) << 10;
  d |= (a_mn_major & 1u) << 15;
  d |= (b_mn_major & 1u) << 16;
  d |= ((n >> 3) & 0x3Fu) << 17;
  d |= ((m >> 4) & 0x1Fu) << 24;
  return d;
}

// ---- split 16-bit representation --------------------------------------------------
// An fp32 value v is carried as two 16-bit floats hi = r16(v), lo = r16(v - hi); products use
// hi*hi + lo*hi + hi*lo with fp32 accumulation.
//   SN_FMT_BF16: 8+8 mantissa bits (~2^-17 relative), fp32 exponent range — used for gradients;
//   SN_FMT_F16 : 11+11 bits (~2^-23 relative for |v| >~ 2^-3, absolute floor 2^-24) — used for
//                activations (O(1) after InstanceNorm) and for weights (pre-scaled by an exact
//                power of two, undone in the epilogue).  fp32-level forward accuracy is what keeps
//                the ReLU / LeakyReLU gates identical to the reference's.
__device__ __forceinline__ void split16(float v, int fmt, uint16_t& hi, uint16_t& lo) {
  if (fmt == SN_FMT_F16) {
    // finite values saturate at the fp16 range (un-normalised layers); NaN / Inf pass through so that a diverged
    // run surfaces as it would in the fp32 reference instead of being masked by fminf / fmaxf
    if (fabsf(v) <= 3.402823466e38f) v = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    hi = __half_as_ushort(h);
    lo = __half_as_ushort(l);
  } else {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    hi = __bfloat16_as_ushort(h);
    lo = __bfloat16_as_ushort(l);
  }
}
__device__ __forceinline__ float decode16(uint16_t x, int fmt) {
  return fmt == SN_FMT_F16 ? __half2float(__ushort_as_half(x)) : __bfloat162float(__ushort_as_bfloat16(x));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#endif  // __CUDACC__

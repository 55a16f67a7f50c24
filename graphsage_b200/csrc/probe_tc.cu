// Developer probe (not in the public header): raw tcgen05.mma issue/execute rate of one CTA per SM with the operand
// layouts K4 uses (K-major SWIZZLE_64B, 8 KB images, M128 x N x K16 bf16), without producers or an epilogue.
//   mode 0: n MMAs back to back, one commit at the end
//   mode 1: a tcgen05.commit (to a barrier nobody waits on) after every 2 MMAs - K4's per-stage commit
//   mode 2: as 1, and the issuing thread also polls an already-completed mbarrier before every 2 MMAs - K4's full wait
//   mode 3: as 0 with the A and B descriptors fixed (same 8 KB images every time: no new shared-memory lines)
#include "tc_common.cuh"

namespace gs {

__device__ unsigned long long g_probe_cycles[4];

__device__ __forceinline__ uint64_t probe_desc64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

template <int N>
__global__ void __launch_bounds__(64, 1) mma_rate_kernel(int mode, int n_mma) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t done_bar, stage_bar, ready_bar;
  __shared__ uint32_t tmem_base_smem;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 26 * 8192 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    mbar_init(&stage_bar, 1);
    mbar_init(&ready_bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 256);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (threadIdx.x == 0) {
    mbar_arrive(&ready_bar);                                  // phase 0 of ready_bar is complete from here on
    constexpr uint32_t idesc = make_idesc(1u, 128, N);
    const uint32_t a0 = smem_u32(smem + 19 * 8192), b0 = smem_u32(smem);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 2) {
      const int st = i >> 1;
      if (mode == 2) mbar_wait(&ready_bar, 0);
      const uint32_t aoff = mode == 3 ? 0u : (uint32_t)(st % 7) * 8192u;
      const uint32_t boff = mode == 3 ? 0u : (uint32_t)(st % (N == 256 ? 9 : 19)) * (N == 256 ? 16384u : 8192u);
      const uint64_t ad = probe_desc64(a0 + aoff), bd = probe_desc64(b0 + boff);
      umma_ss<true>(tmem, ad, bd, idesc, i > 0);
      umma_ss<true>(tmem, ad + 2, bd + 2, idesc, 1u);
      if (mode == 1 || mode == 2) umma_commit(&stage_bar);
    }
    umma_commit(&done_bar);
    const long long t1 = clock64();
    mbar_wait(&done_bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) {
      g_probe_cycles[0] = (unsigned long long)(t1 - t0);     // issue loop
      g_probe_cycles[1] = (unsigned long long)(t2 - t0);     // until the last MMA completed
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// Warp-uniform issue: the WHOLE warp runs the loop with uniform operands and only the instruction itself is predicated
// on elect.sync, so the compiler can keep descriptors in uniform registers (no R2UR / ELECT waterfall per MMA).
//   mode 4: no intermediate commits   mode 5: commit after every 2 MMAs   mode 6: commit after every 4 MMAs
__device__ __forceinline__ void umma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar))
      : "memory");
}

template <int N>
__global__ void __launch_bounds__(32, 1) mma_rate_uniform_kernel(int mode, int n_mma) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t done_bar, stage_bar;
  __shared__ uint32_t tmem_base_smem;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  for (int i = threadIdx.x; i < 26 * 8192 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    mbar_init(&stage_bar, 1);
    fence_mbar_init();
  }
  tmem_alloc(&tmem_base_smem, 256);
  tmem_relinquish();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_smem;
  constexpr uint32_t idesc = make_idesc(1u, 128, N);
  constexpr uint32_t BIMG = N == 256 ? 16384u : 8192u;
  const uint64_t a0 = probe_desc64(smem_u32(smem + 19 * 8192)), b0 = probe_desc64(smem_u32(smem));
  const long long t0 = clock64();
  for (int i = 0; i < n_mma; i += 8) {
    const uint32_t bslot = (uint32_t)(i >> 1) & 7u;            // 8 B images (N = 256) or 16 (N = 128) in turn
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t ad = a0 + (uint64_t)((u * 8192u) >> 4);
      const uint64_t bd = b0 + (uint64_t)((((bslot + u) & (N == 256 ? 7u : 15u)) * BIMG) >> 4);
      umma_elect(tmem, ad, bd, idesc, (i > 0 || u > 0) ? 1u : 0u);
      umma_elect(tmem, ad + 2, bd + 2, idesc, 1u);
      if (mode == 5 || (mode == 6 && (u & 1))) commit_elect(&stage_bar);
    }
  }
  commit_elect(&done_bar);
  const long long t1 = clock64();
  mbar_wait(&done_bar, 0);
  const long long t2 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    g_probe_cycles[0] = (unsigned long long)(t1 - t0);
    g_probe_cycles[1] = (unsigned long long)(t2 - t0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  tmem_dealloc(tmem, 256);
}

}  // namespace gs

extern "C" int32_t gs_debug_mma_rate(int32_t n_cols, int32_t mode, int32_t n_mma, int32_t ctas, unsigned long long* out2) {
  const int smem = 26 * 8192 + 1024;
  if (mode >= 4) {
    if (n_cols == 256) {
      GS_CUDA(cudaFuncSetAttribute(gs::mma_rate_uniform_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      gs::mma_rate_uniform_kernel<256><<<ctas, 32, smem>>>(mode, n_mma);
    } else {
      GS_CUDA(cudaFuncSetAttribute(gs::mma_rate_uniform_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      gs::mma_rate_uniform_kernel<128><<<ctas, 32, smem>>>(mode, n_mma);
    }
  } else if (n_cols == 256) {
    GS_CUDA(cudaFuncSetAttribute(gs::mma_rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gs::mma_rate_kernel<256><<<ctas, 64, smem>>>(mode, n_mma);
  } else {
    GS_CUDA(cudaFuncSetAttribute(gs::mma_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    gs::mma_rate_kernel<128><<<ctas, 64, smem>>>(mode, n_mma);
  }
  GS_CUDA(cudaDeviceSynchronize());
  GS_CUDA(cudaMemcpyFromSymbol(out2, gs::g_probe_cycles, 2 * sizeof(unsigned long long)));
  return GS_OK;
}

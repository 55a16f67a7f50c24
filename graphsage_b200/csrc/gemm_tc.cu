// tcgen05 path of gs_sage_gemm: the neigh_weights / self_weights contraction of the aggregators
// (reference graphsage/aggregators.py:51-64, 110-116, 184-195; Dense graphsage/layers.py:104-116)
// on the 5th-generation tensor cores, fp32 in / fp32 out.
//
//   GS_MATH_TF32X3 : A = A_hi + A_lo, B = B_hi + B_lo (each exactly representable in tf32);
//                    D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi, fp32 accumulate in TMEM  -> fp32-grade result
//   GS_MATH_TF32   : one kind::tf32 pass (operands masked to tf32)
//   GS_MATH_BF16   : one kind::f16 pass, operands rounded to bf16
//
// Structure (one CTA = one 128 x 128 output tile, 320 threads):
//   warps 0-7  A producers: ld.global (coalesced 128-bit, next K-block prefetched in registers) -> split /
//              convert -> st.shared in the UMMA K-major SWIZZLE_128B layout -> fence.proxy.async -> mbarrier.
//              The same warps run the epilogue (tcgen05.ld from TMEM, bias / ReLU, st.global).
//   warp 8     MMA issuer: one elected thread issues tcgen05.mma (SS operands), tcgen05.commit frees stages.
//   warp 9     B loader: cp.async.bulk of pre-swizzled weight tile images (built per call by
//              pack_b_kernel into the caller's workspace) with mbarrier complete_tx.
// A goes through registers on purpose: the hi/lo split (and the bf16 rounding) is arithmetic on the
// operand, and the same producer slot later takes a row-id indirection (gather-A) for the max-pool MLP.
#include "tc_common.cuh"

namespace gs {

constexpr int TC_PRODUCER_WARPS = 8;
constexpr int TC_THREADS = (TC_PRODUCER_WARPS + 2) * 32;

struct TcPart {
  const float* A;
  int64_t lda;
  int32_t K;
  int32_t N;
  int32_t kblocks;     // ceil(K / BK)
  int32_t ntiles;      // ceil(N / 128)
  int64_t img_off;     // byte offset of this part's packed B images in the workspace
  const float* B;
  int64_t ldb;
};

struct TcParams {
  TcPart p[2];
  int32_t n_parts;
  int32_t combine;
  int64_t M;
  const float* bias;
  int32_t act;
  float* out;
  int64_t ldo;
  int32_t tiles_n0;    // number of N tiles of part 0 (CONCAT tile -> part mapping)
  int32_t issue_elect; // 1: warp-uniform elect.sync issue (default), 0: one thread inside `if (lane == 0)`
  const unsigned char* a_img;   // image form (sage_gemm_tc_img_kernel): A operands as tf32 hi/lo tile images (gs_gather_mean_img)
  int32_t a_mtiles;             // 128-row tiles in the A images
  int32_t a_part0;              // A-image part that feeds GEMM part 0 (GEMM part p reads A part a_part0 + p)
};

// ---------------------------------------------------------------------------------------------
// B packing: weights [K, N] row-major fp32 -> per (n tile, k block) tile images of W^T
// (N rows x BK k-elements, K-major, SW128), hi (+ lo) or bf16.  Tiny (<= a few MB), runs per call.
// MODE: 0 = tf32x3 (hi, lo images), 1 = tf32 (hi only), 2 = bf16
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) pack_b_kernel(TcParams prm, unsigned char* __restrict__ ws) {
  constexpr int BK = MODE == 2 ? 64 : 32;
  constexpr int EPC = MODE == 2 ? 8 : 4;        // elements per 16-byte chunk
  constexpr int NIMG = MODE == 0 ? 2 : 1;
  int unit = blockIdx.x;                        // (part, ntile, kblock) flattened
  int pi = 0;
  if (prm.n_parts == 2 && unit >= prm.p[0].ntiles * prm.p[0].kblocks) {
    unit -= prm.p[0].ntiles * prm.p[0].kblocks;
    pi = 1;
  }
  const TcPart& P = prm.p[pi];
  const int nt = unit / P.kblocks, kb = unit % P.kblocks;
  unsigned char* img = ws + P.img_off + ((int64_t)(nt * P.kblocks + kb) * NIMG) * TC_TILE_BYTES;
  for (int q = threadIdx.x; q < 128 * 8; q += blockDim.x) {
    const int c = q >> 7, n = q & 127;           // n fastest: coalesced reads along N
    const int gn = nt * TC_BN + n;
    const int k0 = kb * BK + c * EPC;
    float w[EPC];
#pragma unroll
    for (int e = 0; e < EPC; ++e) w[e] = (gn < P.N && k0 + e < P.K) ? P.B[(int64_t)(k0 + e) * P.ldb + gn] : 0.f;
    const uint32_t off = sw128_off(n, c);
    if constexpr (MODE == 2) {
      __nv_bfloat162 h[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(w[2 * e], w[2 * e + 1]);
      *reinterpret_cast<uint4*>(img + off) = *reinterpret_cast<uint4*>(h);
    } else {
      uint4 hi;
      hi.x = tf32_mask(w[0]); hi.y = tf32_mask(w[1]); hi.z = tf32_mask(w[2]); hi.w = tf32_mask(w[3]);
      *reinterpret_cast<uint4*>(img + off) = hi;
      if constexpr (MODE == 0) {
        uint4 lo;
        lo.x = tf32_mask(w[0] - __uint_as_float(hi.x)); lo.y = tf32_mask(w[1] - __uint_as_float(hi.y));
        lo.z = tf32_mask(w[2] - __uint_as_float(hi.z)); lo.w = tf32_mask(w[3] - __uint_as_float(hi.w));
        *reinterpret_cast<uint4*>(img + TC_TILE_BYTES + off) = lo;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
// timeline probe (CTA (0,0) only): globaltimer stamps at pipeline milestones, read back by gs_debug_read
__device__ unsigned long long g_tc_dbg[32];
__device__ __forceinline__ void dbg_stamp(int slot) {
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_tc_dbg[slot] = t;
  }
}

template <int MODE, bool ASYNC = false>
struct TcCfg {
  static constexpr int BK = MODE == 2 ? 64 : 32;             // K elements per block (128 B of operand row)
  static constexpr int UK = MODE == 2 ? 16 : 8;              // UMMA K per instruction (32 B)
  static constexpr int NIMG = MODE == 0 ? 2 : 1;             // hi (+ lo)
  static constexpr int IMG_BYTES = NIMG * TC_TILE_BYTES;     // one operand's images for one K-block
  // A and B have SEPARATE rings.  A is refilled by the producer warps (their prefetch lives in registers / the
  // raw ring, so few stages suffice).  B tiles come straight from L2 by bulk copy, and a copy can only be posted
  // once its slot is free - so the B ring is deep enough to cover the L2 -> shared latency (NB copies in flight).
  static constexpr int SA = MODE == 0 ? 2 : 3;
  static constexpr int NB = ASYNC ? (MODE == 0 ? 2 : 4) : (MODE == 0 ? 4 : 8);
  static constexpr int RAW_BYTES = TC_BM * BK * 4;
  static constexpr int RAW_SLOTS = ASYNC ? (MODE == 2 ? 3 : 4) : 0;
  static constexpr int B_OFF = SA * IMG_BYTES;
  static constexpr int RAW_OFF = B_OFF + NB * IMG_BYTES;
  static constexpr int SMEM_BYTES = RAW_OFF + RAW_SLOTS * RAW_BYTES + 1024;  // + slack for 1024-B alignment
};


template <int MODE>
__device__ __forceinline__ void load_a_chunk(const TcPart& P, int64_t M, int64_t grow, int gcol, bool vec, float (&v)[8]) {
  constexpr int EPC = MODE == 2 ? 8 : 4;
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  if (grow >= M) return;
  const float* src = P.A + grow * P.lda + gcol;
  if (vec && gcol + EPC <= P.K) {
    float4 a = ldg_nc_f4(reinterpret_cast<const float4*>(src));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    if constexpr (EPC == 8) {
      float4 b = ldg_nc_f4(reinterpret_cast<const float4*>(src) + 1);
      v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
  } else {
#pragma unroll
    for (int e = 0; e < EPC; ++e)
      if (gcol + e < P.K) v[e] = __ldg(src + e);
  }
}

template <int MODE>
__device__ __forceinline__ void store_a_chunk(unsigned char* a_img, int row, int c, const float (&v)[8]) {
  const uint32_t off = sw128_off(row, c);
  if constexpr (MODE == 2) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
    *reinterpret_cast<uint4*>(a_img + off) = *reinterpret_cast<uint4*>(h);
  } else {
    uint4 hi;
    hi.x = tf32_mask(v[0]); hi.y = tf32_mask(v[1]); hi.z = tf32_mask(v[2]); hi.w = tf32_mask(v[3]);
    *reinterpret_cast<uint4*>(a_img + off) = hi;
    if constexpr (MODE == 0) {
      uint4 lo;
      lo.x = tf32_mask(v[0] - __uint_as_float(hi.x)); lo.y = tf32_mask(v[1] - __uint_as_float(hi.y));
      lo.z = tf32_mask(v[2] - __uint_as_float(hi.z)); lo.w = tf32_mask(v[3] - __uint_as_float(hi.w));
      *reinterpret_cast<uint4*>(a_img + TC_TILE_BYTES + off) = lo;
    }
  }
}

template <int MODE, bool ASYNC>
__global__ void __launch_bounds__(TC_THREADS, 1) sage_gemm_tc_kernel(const __grid_constant__ TcParams prm,
                                                                     const unsigned char* __restrict__ ws) {
  using C = TcCfg<MODE, ASYNC>;
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_a[C::SA], empty_a[C::SA], full_b[C::NB], empty_b[C::NB], accum_bar;
  __shared__ uint32_t tmem_base_smem;

  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) dbg_stamp(0);

  // ---- which output tile, which parts feed it
  const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
  int part_lo = 0, part_hi = prm.n_parts, ntile = blockIdx.y, col_off = 0;
  if (prm.combine == GS_COMBINE_CONCAT && prm.n_parts == 2) {
    if ((int)blockIdx.y < prm.tiles_n0) part_hi = 1;
    else { part_lo = 1; ntile = blockIdx.y - prm.tiles_n0; col_off = prm.p[0].N; }
  }
  const int N = prm.p[part_lo].N;
  int total_it = 0;
  for (int pi = part_lo; pi < part_hi; ++pi) total_it += prm.p[pi].kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < C::SA; ++s) {
      mbar_init(&full_a[s], ASYNC ? TC_PRODUCER_WARPS : TC_PRODUCER_WARPS / 2);   // one arrive per producing warp
      mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < C::NB; ++s) {
      mbar_init(&full_b[s], 1);
      mbar_init(&empty_b[s], 1);
    }
    mbar_init(&accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == TC_PRODUCER_WARPS) {   // MMA warp owns the TMEM allocation
    tmem_alloc(&tmem_base_smem, TC_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_smem;
  if (threadIdx.x == 0) dbg_stamp(1);

  if (warp < TC_PRODUCER_WARPS) {
    // =============================== A producers ===============================
    if constexpr (ASYNC) {
      // All 8 warps share every K-block.  Each thread cp.async's ITS OWN 16-byte chunks of the raw fp32 tile
      // into a private position of the staging ring (completion is tracked per thread by cp.async groups, so no
      // cross-thread synchronisation is needed), RAW_SLOTS-1 K-blocks ahead; it then reads them back, splits /
      // converts, and stores the UMMA tile.
      constexpr int EPC = MODE == 2 ? 8 : 4;                 // fp32 elements per UMMA 16-byte chunk
      constexpr int CPT = 4;                                 // UMMA chunks per thread per K-block (1024 / 256)
      constexpr int RAW_PER_CHUNK = EPC / 4;                 // raw 16-byte pieces per UMMA chunk
      unsigned char* raw = smem + C::RAW_OFF;
      const int tid = threadIdx.x;                           // 0..255
      auto locate = [&](int it, int& pi, int& kb) {
        pi = part_lo; kb = it;
        while (kb >= prm.p[pi].kblocks) { kb -= prm.p[pi].kblocks; ++pi; }
      };
      auto issue = [&](int it) {
        if (it < total_it) {
          int pi, kb;
          locate(it, pi, kb);
          const TcPart& P = prm.p[pi];
          unsigned char* slot = raw + (size_t)(it % C::RAW_SLOTS) * C::RAW_BYTES;
#pragma unroll
          for (int i = 0; i < CPT; ++i) {
            const int q = tid + 256 * i;
            const int row = q >> 3, c = q & 7;
            const int64_t grow = m0 + row;
#pragma unroll
            for (int h = 0; h < RAW_PER_CHUNK; ++h) {
              const int gcol = kb * C::BK + c * EPC + h * 4;
              int nbytes = 0;
              if (grow < prm.M && gcol < P.K) nbytes = min(4, P.K - gcol) * 4;
              const float* src = nbytes ? (P.A + grow * P.lda + gcol) : P.A;   // always a valid, 16-B aligned address
              cp_async16(slot + (size_t)(q * RAW_PER_CHUNK + h) * 16, src, nbytes);
            }
          }
        }
        cp_async_commit();                                   // empty groups keep the group count uniform
      };
#pragma unroll
      for (int j = 0; j < C::RAW_SLOTS - 1; ++j) issue(j);
      for (int it = 0; it < total_it; ++it) {
        issue(it + C::RAW_SLOTS - 1);
        cp_async_wait<C::RAW_SLOTS - 1>();                   // this thread's pieces of K-block `it` have landed
        const unsigned char* slot = raw + (size_t)(it % C::RAW_SLOTS) * C::RAW_BYTES;
        float v[CPT][8];
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          const int q = tid + 256 * i;
#pragma unroll
          for (int h = 0; h < RAW_PER_CHUNK; ++h) {
            const float4 f = *reinterpret_cast<const float4*>(slot + (size_t)(q * RAW_PER_CHUNK + h) * 16);
            v[i][h * 4 + 0] = f.x; v[i][h * 4 + 1] = f.y; v[i][h * 4 + 2] = f.z; v[i][h * 4 + 3] = f.w;
          }
        }
        const int s = it % C::SA;
        const uint32_t ph = (uint32_t)(it / C::SA) & 1u;
        mbar_wait(&empty_a[s], ph ^ 1u);
        unsigned char* a_img = smem + (size_t)s * C::IMG_BYTES;
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          const int q = tid + 256 * i;
          store_a_chunk<MODE>(a_img, q >> 3, q & 7, v[i]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[s]);
      }
    } else {
    // two groups of 4 warps alternate K-blocks; each thread owns 8 chunks (rows tg>>3 + 16 i, chunk tg & 7)
    const int group = warp >> 2;
    const int tg = threadIdx.x & 127;
    const int c = tg & 7, r0 = tg >> 3;
    float cur[8][8];
    auto locate = [&](int it, int& pi, int& kb) {
      pi = part_lo; kb = it;
      while (kb >= prm.p[pi].kblocks) { kb -= prm.p[pi].kblocks; ++pi; }
    };
    auto fetch = [&](int it, float (&dst)[8][8]) {
      int pi, kb;
      locate(it, pi, kb);
      const TcPart& P = prm.p[pi];
      const bool vec = ((P.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.A) & 15u) == 0);
      const int gcol = kb * C::BK + c * (MODE == 2 ? 8 : 4);
#pragma unroll
      for (int i = 0; i < 8; ++i) load_a_chunk<MODE>(P, prm.M, m0 + r0 + 16 * i, gcol, vec, dst[i]);
    };
    int it = group;
    if (it < total_it) fetch(it, cur);
    for (; it < total_it; it += 2) {
      const int s = it % C::SA;
      const uint32_t ph = (uint32_t)(it / C::SA) & 1u;
      float nxt[8][8];
      const bool more = it + 2 < total_it;
      if (more) fetch(it + 2, nxt);                    // prefetch this group's next K-block
      mbar_wait(&empty_a[s], ph ^ 1u);
      unsigned char* a_img = smem + (size_t)s * C::IMG_BYTES;
#pragma unroll
      for (int i = 0; i < 8; ++i) store_a_chunk<MODE>(a_img, r0 + 16 * i, c, cur[i]);
      fence_proxy_async();                             // generic-proxy stores -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_a[s]);
      if (threadIdx.x == 0 && it == 0) dbg_stamp(2);
      if (threadIdx.x == 0 && it + 2 >= total_it) dbg_stamp(3);
      if (more) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int e = 0; e < 8; ++e) cur[i][e] = nxt[i][e];
      }
    }
    }
    // =============================== epilogue ===============================
    if (threadIdx.x == 0) dbg_stamp(4);
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    if (threadIdx.x == 0) dbg_stamp(5);
    const int q = warp & 3;                           // TMEM lane quarter this warp may touch
    const int half = warp >> 2;                       // columns [64*half, 64*half + 64)
    const int64_t grow = m0 + q * 32 + lane;
#pragma unroll
    for (int cb = 0; cb < 2; ++cb) {
      const int col0 = half * 64 + cb * 32;
      uint32_t r[32];
      tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, r);
      tmem_ld_wait();
      if (grow < prm.M) {
        const int gn0 = ntile * TC_BN + col0;
        float* dst = prm.out + grow * prm.ldo + col_off + gn0;
        const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && gn0 + 32 <= N;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[e] = __uint_as_float(r[j + e]);
            if (prm.bias && gn0 + j + e < N) v[e] += prm.bias[col_off + gn0 + j + e];
            if (prm.act == GS_ACT_RELU) v[e] = fmaxf(v[e], 0.f);
          }
          if (vec) {
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (gn0 + j + e < N) dst[j + e] = v[e];
          }
        }
      }
    }
    tc_fence_before();
    if (threadIdx.x == 0) dbg_stamp(6);
  } else if (warp == TC_PRODUCER_WARPS) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc = make_idesc(MODE == 2 ? 1u : 2u, TC_BM, TC_BN);
    for (int it = 0; it < total_it; ++it) {
      const int s = it % C::SA, sb = it % C::NB;
      mbar_wait(&full_a[s], (uint32_t)(it / C::SA) & 1u);
      if (lane == 0 && it == 0) dbg_stamp(8);
      mbar_wait(&full_b[sb], (uint32_t)(it / C::NB) & 1u);
      tc_fence_after();
      if (lane == 0 && it == 0) dbg_stamp(9);
      if (lane == 0 && it == 1) dbg_stamp(10);
      if (lane == 0 && it == 8) dbg_stamp(11);
      if (lane == 0 && it == total_it - 1) dbg_stamp(12);
      const uint32_t a_base = smem_u32(smem + (size_t)s * C::IMG_BYTES);
      const uint32_t b_base = smem_u32(smem + C::B_OFF + (size_t)sb * C::IMG_BYTES);
      const uint64_t a_hi = make_smem_desc(a_base), b_hi = make_smem_desc(b_base);
      if (prm.issue_elect) {
        // whole warp, uniform operands, elect.sync on the instruction (tc_common.cuh: umma_ss_elect)
#pragma unroll
        for (int k = 0; k < C::BK / C::UK; ++k) {
          const uint64_t koff = (uint64_t)((k * 32) >> 4);          // 32 B per UMMA K step inside the swizzle atom
          const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
          umma_ss_elect<MODE == 2>(tmem_acc, a_hi + koff, b_hi + koff, idesc, acc);
          if constexpr (MODE == 0) {
            const uint64_t a_lo = make_smem_desc(a_base + TC_TILE_BYTES), b_lo = make_smem_desc(b_base + TC_TILE_BYTES);
            umma_ss_elect<false>(tmem_acc, a_hi + koff, b_lo + koff, idesc, 1u);
            umma_ss_elect<false>(tmem_acc, a_lo + koff, b_hi + koff, idesc, 1u);
          }
        }
        umma_commit_elect(&empty_a[s]);                                // A stage and B slot reusable once these MMAs retire
        umma_commit_elect(&empty_b[sb]);
        if (it == total_it - 1) umma_commit_elect(&accum_bar);         // accumulator complete
        if (lane == 0 && it == total_it - 1) dbg_stamp(13);
      } else if (lane == 0) {
#pragma unroll
        for (int k = 0; k < C::BK / C::UK; ++k) {
          const uint64_t koff = (uint64_t)((k * 32) >> 4);
          const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
          umma_ss<MODE == 2>(tmem_acc, a_hi + koff, b_hi + koff, idesc, acc);
          if constexpr (MODE == 0) {
            const uint64_t a_lo = make_smem_desc(a_base + TC_TILE_BYTES), b_lo = make_smem_desc(b_base + TC_TILE_BYTES);
            umma_ss<false>(tmem_acc, a_hi + koff, b_lo + koff, idesc, 1u);
            umma_ss<false>(tmem_acc, a_lo + koff, b_hi + koff, idesc, 1u);
          }
        }
        umma_commit(&empty_a[s]);
        umma_commit(&empty_b[sb]);
        if (it == total_it - 1) umma_commit(&accum_bar);
        if (it == total_it - 1) dbg_stamp(13);
      }
      __syncwarp();
    }
  } else {
    // =============================== B loader ===============================
    if (lane == 0) {
      int it = 0;
      for (int pi = part_lo; pi < part_hi; ++pi) {
        const TcPart& P = prm.p[pi];
        const unsigned char* base = ws + P.img_off + (int64_t)ntile * P.kblocks * C::NIMG * TC_TILE_BYTES;
        for (int kb = 0; kb < P.kblocks; ++kb, ++it) {
          const int sb = it % C::NB;
          const uint32_t ph = (uint32_t)(it / C::NB) & 1u;
          mbar_wait(&empty_b[sb], ph ^ 1u);
          mbar_expect_tx(&full_b[sb], C::IMG_BYTES);
          bulk_g2s(smem + C::B_OFF + (size_t)sb * C::IMG_BYTES, base + (int64_t)kb * C::IMG_BYTES, C::IMG_BYTES,
                   &full_b[sb]);
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (threadIdx.x == 0) dbg_stamp(7);
  if (warp == TC_PRODUCER_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, TC_BN);
  }
}

// ---------------------------------------------------------------------------------------------
// Image form (tf32x3 only): the A operand arrives as the tf32 hi / lo UMMA tile images the fused gather wrote
// (gs_gather_mean_img), so a K-block of A is ONE 32 KB bulk copy - no producer warps, no register round trip, no hi/lo
// split here, no generic->async proxy fence.  Same arithmetic as sage_gemm_tc_kernel<0> (A_hi*B_hi + A_hi*B_lo +
// A_lo*B_hi in that order, fp32 accumulate in TMEM): results are bit-identical to it.
//   warp 0: A loader (bulk copies)   warp 1: B loader   warp 2: MMA issuer (+ TMEM alloc)   warps 4-7: epilogue
// Three 32 KB stages per operand (192 KB): the ring covers the L2 -> shared latency, which the register-staged form could not.
// ---------------------------------------------------------------------------------------------
constexpr int TCI_THREADS = 256;
constexpr int TCI_IMG = 2 * TC_TILE_BYTES;      // hi + lo image of one K-block
constexpr int TCI_SA = 3, TCI_NB = 3;
constexpr int TCI_SMEM = (TCI_SA + TCI_NB) * TCI_IMG + 1024;

__global__ void __launch_bounds__(TCI_THREADS, 1) sage_gemm_tc_img_kernel(const __grid_constant__ TcParams prm,
                                                                          const unsigned char* __restrict__ ws) {
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_a[TCI_SA], empty_a[TCI_SA], full_b[TCI_NB], empty_b[TCI_NB], accum_bar;
  __shared__ uint32_t tmem_base_smem;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* a_ring = smem;
  unsigned char* b_ring = smem + TCI_SA * TCI_IMG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
  int part_lo = 0, part_hi = prm.n_parts, ntile = blockIdx.y, col_off = 0;
  if (prm.combine == GS_COMBINE_CONCAT && prm.n_parts == 2) {
    if ((int)blockIdx.y < prm.tiles_n0) part_hi = 1;
    else { part_lo = 1; ntile = blockIdx.y - prm.tiles_n0; col_off = prm.p[0].N; }
  }
  const int N = prm.p[part_lo].N;
  int total_it = 0;
  for (int pi = part_lo; pi < part_hi; ++pi) total_it += prm.p[pi].kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TCI_SA; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < TCI_NB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    mbar_init(&accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(&tmem_base_smem, TC_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = tmem_base_smem;

  if (warp == 0) {
    // =============================== A loader ===============================
    if (lane == 0) {
      int it = 0;
      for (int pi = part_lo; pi < part_hi; ++pi) {
        const TcPart& P = prm.p[pi];
        const unsigned char* base = prm.a_img + (((int64_t)(prm.a_part0 + pi) * prm.a_mtiles + blockIdx.x) * P.kblocks) * TCI_IMG;
        for (int kb = 0; kb < P.kblocks; ++kb, ++it) {
          const int sa = it % TCI_SA;
          mbar_wait(&empty_a[sa], ((uint32_t)(it / TCI_SA) & 1u) ^ 1u);
          mbar_expect_tx(&full_a[sa], TCI_IMG);
          bulk_g2s(a_ring + (size_t)sa * TCI_IMG, base + (int64_t)kb * TCI_IMG, TCI_IMG, &full_a[sa]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =============================== B loader ===============================
    if (lane == 0) {
      int it = 0;
      for (int pi = part_lo; pi < part_hi; ++pi) {
        const TcPart& P = prm.p[pi];
        const unsigned char* base = ws + P.img_off + (int64_t)ntile * P.kblocks * TCI_IMG;
        for (int kb = 0; kb < P.kblocks; ++kb, ++it) {
          const int sb = it % TCI_NB;
          mbar_wait(&empty_b[sb], ((uint32_t)(it / TCI_NB) & 1u) ^ 1u);
          mbar_expect_tx(&full_b[sb], TCI_IMG);
          bulk_g2s(b_ring + (size_t)sb * TCI_IMG, base + (int64_t)kb * TCI_IMG, TCI_IMG, &full_b[sb]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 2) {
    // =============================== MMA issuer ===============================
    constexpr uint32_t idesc = make_idesc(2u, TC_BM, TC_BN);            // tf32 x tf32 -> fp32
    for (int it = 0; it < total_it; ++it) {
      const int sa = it % TCI_SA, sb = it % TCI_NB;
      mbar_wait_uniform(&full_a[sa], (uint32_t)(it / TCI_SA) & 1u);
      mbar_wait_uniform(&full_b[sb], (uint32_t)(it / TCI_NB) & 1u);
      tc_fence_after();
      const uint32_t a_base = smem_u32(a_ring + (size_t)sa * TCI_IMG), b_base = smem_u32(b_ring + (size_t)sb * TCI_IMG);
      const uint64_t a_hi = make_smem_desc(a_base), b_hi = make_smem_desc(b_base);
      const uint64_t a_lo = make_smem_desc(a_base + TC_TILE_BYTES), b_lo = make_smem_desc(b_base + TC_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < 4; ++k) {                                     // four K = 8 steps per 32-column K-block
        const uint64_t koff = (uint64_t)((k * 32) >> 4);
        umma_ss_elect<false>(tmem_acc, a_hi + koff, b_hi + koff, idesc, (it > 0 || k > 0) ? 1u : 0u);
        umma_ss_elect<false>(tmem_acc, a_hi + koff, b_lo + koff, idesc, 1u);
        umma_ss_elect<false>(tmem_acc, a_lo + koff, b_hi + koff, idesc, 1u);
      }
      umma_commit_elect(&empty_a[sa]);
      umma_commit_elect(&empty_b[sb]);
      if (it == total_it - 1) umma_commit_elect(&accum_bar);
    }
  } else if (warp >= 4) {
    // =============================== epilogue ===============================
    mbar_wait(&accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3;                           // TMEM lane quarter of this warp
    const int64_t grow = m0 + q * 32 + lane;
#pragma unroll 1
    for (int cb = 0; cb < 4; ++cb) {
      const int col0 = cb * 32;
      uint32_t r[32];
      tmem_ld_32x32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, r);
      tmem_ld_wait();
      if (grow < prm.M) {
        const int gn0 = ntile * TC_BN + col0;
        float* dst = prm.out + grow * prm.ldo + col_off + gn0;
        const bool vec = ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) && gn0 + 32 <= N;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[e] = __uint_as_float(r[j + e]);
            if (prm.bias && gn0 + j + e < N) v[e] += prm.bias[col_off + gn0 + j + e];
            if (prm.act == GS_ACT_RELU) v[e] = fmaxf(v[e], 0.f);
          }
          if (vec) {
            *reinterpret_cast<float4*>(dst + j) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (gn0 + j + e < N) dst[j + e] = v[e];
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, TC_BN);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int mode_of(int32_t math) { return math == GS_MATH_TF32X3 ? 0 : math == GS_MATH_TF32 ? 1 : 2; }

static void fill_parts(TcParams& prm, int64_t M, const gs_gemm_part* parts, int32_t n_parts, int mode) {
  const int BK = mode == 2 ? 64 : 32;
  const int nimg = mode == 0 ? 2 : 1;
  memset(&prm, 0, sizeof(prm));
  prm.n_parts = n_parts;
  prm.M = M;
  int64_t off = 0;
  for (int i = 0; i < n_parts; ++i) {
    TcPart& P = prm.p[i];
    P.A = parts[i].A; P.lda = parts[i].lda; P.K = parts[i].K; P.N = parts[i].N;
    P.B = parts[i].B; P.ldb = parts[i].ldb;
    P.kblocks = (P.K + BK - 1) / BK;
    P.ntiles = (P.N + TC_BN - 1) / TC_BN;
    P.img_off = off;
    off += (int64_t)P.kblocks * P.ntiles * nimg * TC_TILE_BYTES;
  }
  prm.tiles_n0 = prm.p[0].ntiles;
  prm.issue_elect = tuning("mma_issue", 1) != 0;
}

int64_t sage_gemm_tc_workspace(int64_t M, const gs_gemm_part* parts, int32_t n_parts, int32_t math) {
  TcParams prm;
  fill_parts(prm, M, parts, n_parts, mode_of(math));
  const TcPart& L = prm.p[n_parts - 1];
  const int nimg = mode_of(math) == 0 ? 2 : 1;
  return L.img_off + (int64_t)L.kblocks * L.ntiles * nimg * TC_TILE_BYTES;
}

template <int MODE>
static int32_t launch_pack(const TcParams& prm, unsigned char* ws, cudaStream_t st) {
  int units = prm.p[0].ntiles * prm.p[0].kblocks + (prm.n_parts == 2 ? prm.p[1].ntiles * prm.p[1].kblocks : 0);
  pack_b_kernel<MODE><<<units, 256, 0, st>>>(prm, ws);
  return launch_check("pack_b_kernel");
}

template <int MODE, bool ASYNC>
static int32_t launch_tc_impl(const TcParams& prm, const unsigned char* ws, cudaStream_t st) {
  using C = TcCfg<MODE, ASYNC>;
  {
    const int32_t rc_attr = ensure_dyn_smem((const void*)sage_gemm_tc_kernel<MODE, ASYNC>, C::SMEM_BYTES);
    if (rc_attr != GS_OK) return rc_attr;
  }
  int tiles_n = prm.p[0].ntiles;
  if (prm.combine == GS_COMBINE_CONCAT && prm.n_parts == 2) tiles_n += prm.p[1].ntiles;
  dim3 grid((unsigned)((prm.M + TC_BM - 1) / TC_BM), (unsigned)tiles_n);
  sage_gemm_tc_kernel<MODE, ASYNC><<<grid, TC_THREADS, C::SMEM_BYTES, st>>>(prm, ws);
  return launch_check("sage_gemm_tc_kernel");
}

template <int MODE>
static int32_t launch_tc(const TcParams& prm, const unsigned char* ws, cudaStream_t st) {
  bool aligned = true;                         // cp.async needs 16-byte aligned rows
  for (int i = 0; i < prm.n_parts; ++i)
    aligned = aligned && (prm.p[i].lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(prm.p[i].A) & 15u) == 0);
  if (aligned && tuning("gemm_async", 0)) return launch_tc_impl<MODE, true>(prm, ws, st);
  return launch_tc_impl<MODE, false>(prm, ws, st);
}

int32_t tc_debug_read(unsigned long long* out_host, int n) {
  if (n > 32) n = 32;
  GS_CUDA(cudaDeviceSynchronize());
  GS_CUDA(cudaMemcpyFromSymbol(out_host, g_tc_dbg, sizeof(unsigned long long) * n));
  return GS_OK;
}

int32_t sage_gemm_tc_pack(const gs_gemm_part* parts, int32_t n_parts, int32_t math, void* workspace, cudaStream_t st) {
  GS_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 127u) == 0,
             "gs_sage_gemm_pack: workspace must be non-NULL and 128-byte aligned");
  const int mode = mode_of(math);
  TcParams prm;
  fill_parts(prm, 0, parts, n_parts, mode);
  unsigned char* ws = (unsigned char*)workspace;
  if (mode == 0) return launch_pack<0>(prm, ws, st);
  if (mode == 1) return launch_pack<1>(prm, ws, st);
  return launch_pack<2>(prm, ws, st);
}

int32_t sage_gemm_tc_img(int64_t M, const gs_gemm_part* parts, int32_t n_parts, int32_t combine, const float* bias, int32_t act,
                         float* out, int64_t ldo, const void* workspace, const void* a_images, int32_t a_part0,
                         cudaStream_t st) {
  GS_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 127u) == 0,
             "gs_sage_gemm_img: packed weights missing or not 128-byte aligned");
  GS_REQUIRE(a_images != nullptr && (reinterpret_cast<uintptr_t>(a_images) & 1023u) == 0,
             "gs_sage_gemm_img: A images missing or not 1024-byte aligned");
  TcParams prm;
  fill_parts(prm, M, parts, n_parts, 0);
  for (int i = 1; i < n_parts; ++i)
    GS_REQUIRE(prm.p[i].K == prm.p[0].K, "gs_sage_gemm_img: every part must have the K of the gathered rows");
  prm.combine = combine; prm.bias = bias; prm.act = act; prm.out = out; prm.ldo = ldo;
  prm.a_img = (const unsigned char*)a_images;
  prm.a_mtiles = (int32_t)((M + TC_BM - 1) / TC_BM);
  prm.a_part0 = a_part0;
  const int32_t rc_attr = ensure_dyn_smem((const void*)sage_gemm_tc_img_kernel, TCI_SMEM);
  if (rc_attr != GS_OK) return rc_attr;
  int tiles_n = prm.p[0].ntiles;
  if (combine == GS_COMBINE_CONCAT && n_parts == 2) tiles_n += prm.p[1].ntiles;
  dim3 grid((unsigned)prm.a_mtiles, (unsigned)tiles_n);
  sage_gemm_tc_img_kernel<<<grid, TCI_THREADS, TCI_SMEM, st>>>(prm, (const unsigned char*)workspace);
  return launch_check("sage_gemm_tc_img_kernel");
}

int32_t sage_gemm_tc(int64_t M, const gs_gemm_part* parts, int32_t n_parts, int32_t combine, const float* bias,
                     int32_t act, int32_t math, float* out, int64_t ldo, const void* workspace, cudaStream_t st) {
  GS_REQUIRE(workspace != nullptr, "gs_sage_gemm: tensor-core math modes need the workspace (gs_sage_gemm_workspace_bytes)");
  GS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 127u) == 0, "gs_sage_gemm: workspace must be 128-byte aligned");
  const int mode = mode_of(math);
  TcParams prm;
  fill_parts(prm, M, parts, n_parts, mode);
  prm.combine = combine;
  prm.bias = bias;
  prm.act = act;
  prm.out = out;
  prm.ldo = ldo;
  const unsigned char* ws = (const unsigned char*)workspace;
  if (mode == 0) return launch_tc<0>(prm, ws, st);
  if (mode == 1) return launch_tc<1>(prm, ws, st);
  return launch_tc<2>(prm, ws, st);
}

}  // namespace gs

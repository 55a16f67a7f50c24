// K4: the max-pool aggregator's neighbour branch as ONE kernel on tcgen05 (bf16 operands, fp32 accumulate):
//   out[g, h] = max_{j<k} relu( table[row(g, j), :] . Wm[:, h] + bm[h] )
//   reference graphsage/aggregators.py:176-182 (reshape -> Dense(relu, bias) -> reshape -> reduce_max) with
//   graphsage/layers.py:104-116 and the feature gather of graphsage/models.py:299 fused in front of it:
//   neither the gathered [n*k, F] rows nor the [n*k, hidden] MLP activations ever touch HBM.
//
// Persistent CTAs.  A CTA owns one 128-wide slice of the hidden dimension for its whole life: that slice of
// Wm^T (<= 10 K-blocks x 16 KB, pre-swizzled bf16 tile images) is bulk-copied into shared memory ONCE and stays
// resident.  The CTA then walks M tiles (G = floor(128 / k) whole fanout groups per tile, zero rows after them):
//   warps 0-3  gather-A producers: cp.async 16-byte pieces of the addressed table rows straight into the UMMA
//              K-major SWIZZLE_128B tile (bf16 in the table = bf16 in the tile: no conversion), 2 K-blocks ahead.
//              (Measured alternatives - 3-5 K-blocks of 128-bit register loads per thread, with 4 or 8 producer
//              warps, and an L2 prefetch of the next tile - were slower; the limiter is shared-memory capacity:
//              the 160 KB resident weight slice leaves 48 KB for A stages.  See DESIGN.md section 4.)
//   warp  4    MMA issuer: tcgen05.mma.kind::f16 (M128 N128 K16), accumulators double-buffered in TMEM
//   warp  5    loads the resident weight slice (cp.async.bulk + mbarrier)
//   warps 6-9  epilogue: tcgen05.ld -> + bias -> ReLU -> staged transpose in shared memory -> max over the k rows
//              of each group -> coalesced store; overlaps the next tile's MMAs (second TMEM buffer)
#include <cuda.h>   // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint (no libcuda link)

#include "tc_common.cuh"

namespace gs {

// K-blocks are 32 bf16 columns = 64-byte operand rows (UMMA K-major SWIZZLE_64B): an 8 KB stage instead of 16 KB, so
// six A stages fit beside the resident weights (with 16 KB stages only three did, and the stage a producer needed
// next was always the one whose MMA had just been published - a structural bubble), and K pads to 32, not 64.
constexpr int MP_KCOLS = 32;                  // bf16 columns per K-block
constexpr int MP_IMG = TC_BM * 64;            // one operand K-block image: 128 rows x 64 B
constexpr int MP_MAX_KB = 20;                 // resident weight K-blocks (K <= 640)
constexpr int MP_RING = 26;                   // 8 KB slots shared by the resident weights (kblocks) and the A stages
constexpr int MP_PROD_WARPS = 4;               // gather-A producer warps
constexpr int MP_THREADS = (MP_PROD_WARPS + 6) * 32;
constexpr int MP_STAGE_LD = 129;              // padded row length of the epilogue staging [32][129]
constexpr int MP_SMEM = MP_RING * MP_IMG + 32 * MP_STAGE_LD * 4 + 1024;

struct MpParams {
  const __nv_bfloat16* table;   // [n_rows, pitch]
  int64_t n_rows, pitch;
  int32_t K, kblocks;
  const int32_t* row_ids;       // [n_groups * k] or NULL
  int64_t row0;                 // used when row_ids == NULL: row(g, j) = row0 + g*k + j
  int64_t n_groups;
  int32_t k, G;                 // fanout, groups per tile
  int64_t n_tiles;
  int32_t hidden, n_slices;
  const unsigned char* wimg;    // packed Wm^T: [n_slices][kblocks][16 KB]
  const float* bias;            // [hidden] or NULL
  int32_t pool_mean;            // 0: max over the fanout (MaxPoolingAggregator), 1: mean (MeanPoolingAggregator)
  float* out;                   // [n_groups, hidden]
  int64_t ldo;
  int32_t issue_elect;          // 1: warp-uniform elect.sync issue (default), 0: one thread inside `if (lane == 0)`
};

// byte offset of 16-byte chunk c (0..3) of row r inside a K-major SWIZZLE_64B image (Swizzle<2,4,3>: address bits
// [4,6) ^= bits [7,9); rows are 64 B apart, so bits [7,9) = (r >> 1) & 3)
__host__ __device__ __forceinline__ uint32_t sw64_off(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

// K-major SWIZZLE_64B shared-memory matrix descriptor: SBO = 8 rows x 64 B = 512 B, layout type 4
__device__ __forceinline__ uint64_t make_smem_desc64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

// Wm [K, hidden] row-major fp32 -> bf16 tile images of Wm^T (128 hidden rows x 32 k, K-major, SW64)
__global__ void __launch_bounds__(256) maxpool_pack_kernel(const float* __restrict__ W, int64_t ldw, int K, int hidden,
                                                           int kblocks, unsigned char* __restrict__ img) {
  const int slice = blockIdx.x / kblocks, kb = blockIdx.x % kblocks;
  unsigned char* dst = img + ((int64_t)slice * kblocks + kb) * MP_IMG;
  for (int q = threadIdx.x; q < 128 * 4; q += blockDim.x) {
    const int c = q >> 7, n = q & 127;
    const int gn = slice * 128 + n, k0 = kb * MP_KCOLS + c * 8;
    __nv_bfloat162 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float a = (gn < hidden && k0 + 2 * e < K) ? W[(int64_t)(k0 + 2 * e) * ldw + gn] : 0.f;
      float b = (gn < hidden && k0 + 2 * e + 1 < K) ? W[(int64_t)(k0 + 2 * e + 1) * ldw + gn] : 0.f;
      h[e] = __floats2bfloat162_rn(a, b);
    }
    *reinterpret_cast<uint4*>(dst + sw64_off(n, c)) = *reinterpret_cast<uint4*>(h);
  }
}

// timeline probe (CTA 0): globaltimer stamps, 8 slots per tile for the first 12 tiles
__device__ unsigned long long g_mp_dbg[128];
__device__ __forceinline__ void mp_stamp(uint32_t tcount, int slot) {
  if (blockIdx.x == 0 && tcount < 12) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_mp_dbg[tcount * 8 + slot] = t;
  }
}

// Optional epilogue cycle breakdown (build with GS_EXTRA_NVCC_FLAGS=-DGS_K4_EPI_PROBE; compiled out otherwise):
// CTA 0, epilogue thread 0 accumulates clock64 deltas into g_mp_dbg[100..105] =
// {tmem ld + wait, staging stores, barrier 1, pooling loop + global stores, barrier 2, column blocks counted}
#ifdef GS_K4_EPI_PROBE
#define MP_EPI_T(var) const long long var = clock64()
#define MP_EPI_ACC(slot, a, b) \
  do {                         \
    if (blockIdx.x == 0 && et == 0) g_mp_dbg[100 + (slot)] += (unsigned long long)((b) - (a)); \
  } while (0)
#else
#define MP_EPI_T(var) \
  do {                \
  } while (0)
#define MP_EPI_ACC(slot, a, b) \
  do {                         \
  } while (0)
#endif

// ---- roles shared by the kernel variants (forced inline: one copy of the logic, no call overhead) ----

// MMA issuer warp: waits for each A stage, issues the two K = 16 MMAs of the K-block against the resident weights,
// commits the stage back to the producers and, after a tile's last K-block, the accumulator to the epilogue.
template <int MP_SA, int CL = 1>
__device__ __forceinline__ void mp_mma_role(const MpParams& prm, int lane, int kblocks, int64_t tile0, int64_t tile_step,
                                            uint32_t tmem_base, uint64_t* full_a, uint64_t* empty_a, uint64_t* acc_full,
                                            uint64_t* acc_empty, uint64_t& b_full, unsigned char* a_ring,
                                            unsigned char* b_res) {
  // =============================== MMA issuer ===============================
  constexpr uint32_t idesc = make_idesc(1u, TC_BM, TC_BN);          // bf16 x bf16 -> fp32
  mbar_wait(&b_full, 0);
  uint32_t it = 0, tcount = 0;
  for (int64_t t = tile0; t < prm.n_tiles; t += tile_step, ++tcount) {
    const uint32_t buf = tcount & 1u;
    if (lane == 0) mp_stamp(tcount, 2);
    mbar_wait(&acc_empty[buf], ((tcount >> 1) & 1u) ^ 1u);          // epilogue has drained this accumulator
    tc_fence_after();
    if (lane == 0) mp_stamp(tcount, 3);
    const uint32_t tmem_acc = tmem_base + buf * 128u;
    for (int kb = 0; kb < kblocks; ++kb, ++it) {
      const int s = it % MP_SA;
      mbar_wait(&full_a[s], (it / MP_SA) & 1u);
      tc_fence_after();
      const uint64_t adesc = make_smem_desc64(smem_u32(a_ring + (size_t)s * MP_IMG));
      const uint64_t bdesc = make_smem_desc64(smem_u32(b_res + (size_t)kb * MP_IMG));
      if (prm.issue_elect) {
        // whole warp, uniform operands, elect.sync on the instruction (tc_common.cuh: umma_ss_elect)
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2)            // two K = 16 steps per 32-column K-block (32 B apart inside the atom)
          umma_ss_elect<true>(tmem_acc, adesc + (uint64_t)(k2 * 2), bdesc + (uint64_t)(k2 * 2), idesc,
                              (kb > 0 || k2 > 0) ? 1u : 0u);
        if constexpr (CL > 1)                       // frees the stage in every CTA of the cluster
          umma_commit_elect_multicast(&empty_a[s], (uint16_t)((1u << CL) - 1u));
        else
          umma_commit_elect(&empty_a[s]);
        if (kb == kblocks - 1) umma_commit_elect(&acc_full[buf]);
        if (lane == 0 && kb == kblocks - 1) mp_stamp(tcount, 4);
      } else if (lane == 0) {
#pragma unroll
        for (int k2 = 0; k2 < 2; ++k2)
          umma_ss<true>(tmem_acc, adesc + (uint64_t)(k2 * 2), bdesc + (uint64_t)(k2 * 2), idesc, (kb > 0 || k2 > 0) ? 1u : 0u);
        umma_commit(&empty_a[s]);
        if (kb == kblocks - 1) umma_commit(&acc_full[buf]);
        if (kb == kblocks - 1) mp_stamp(tcount, 4);
      }
      __syncwarp();
    }
  }
}

// Weight loader warp: one bulk copy per K-block image of this CTA's slice, once.
__device__ __forceinline__ void mp_weights_role(const MpParams& prm, int lane, int slice, int kblocks, uint64_t& b_full,
                                                unsigned char* b_res) {
  // =============================== resident weight slice ===============================
  if (lane == 0) {
    mbar_expect_tx(&b_full, (uint32_t)(kblocks * MP_IMG));
    const unsigned char* src = prm.wimg + (int64_t)slice * kblocks * MP_IMG;
    for (int kb = 0; kb < kblocks; ++kb)
      bulk_g2s(b_res + (size_t)kb * MP_IMG, src + (int64_t)kb * MP_IMG, MP_IMG, &b_full);
  }
  __syncwarp();
}

// Epilogue (four consecutive warps; et = 0..127 is the thread's index among them).
__device__ __forceinline__ void mp_epilogue_role(const MpParams& prm, int et, int warp, int lane, int slice, int64_t tile0,
                                                 int64_t tile_step, uint32_t tmem_base, uint64_t* acc_full,
                                                 uint64_t* acc_empty, float* stage, float* bias_s) {
  // =============================== epilogue ===============================
  const int q = warp & 3;                         // TMEM lane quarter of this warp (four consecutive warps cover 0..3)
  const int row = q * 32 + lane;                  // tile row held by this thread
  const int k = prm.k, G = prm.G;
  bias_s[et] = prm.bias ? prm.bias[slice * 128 + et] : 0.f;     // this CTA's 128 bias values, once
  named_bar_sync(1, 128);
  uint32_t tcount = 0;
  for (int64_t t = tile0; t < prm.n_tiles; t += tile_step, ++tcount) {
    const uint32_t buf = tcount & 1u;
    if (et == 0) mp_stamp(tcount, 5);
    mbar_wait(&acc_full[buf], (tcount >> 1) & 1u);
    tc_fence_after();
    if (et == 0) mp_stamp(tcount, 6);
    const uint32_t tmem_acc = tmem_base + buf * 128u + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int cb = 0; cb < 4; ++cb) {
      uint32_t r[32];
      MP_EPI_T(e0);
      tmem_ld_32x32(tmem_acc + (uint32_t)(cb * 32), r);
      tmem_ld_wait();
      MP_EPI_T(e1);
      const int hcol0 = slice * 128 + cb * 32;
      // raw accumulators go to the staging tile; bias and ReLU are applied AFTER the max
      // (max_j relu(x_j + b) == relu(max_j x_j + b): b is per column, relu is monotone)
#pragma unroll
      for (int j = 0; j < 32; ++j) stage[j * MP_STAGE_LD + row] = __uint_as_float(r[j]);
      MP_EPI_T(e2);
      named_bar_sync(1, 128);
      MP_EPI_T(e3);
      // thread = (column cc, group residue): max over each group's k consecutive rows, 8 independent
      // shared loads per batch
      {
        const int cc = et & 31;
        for (int g = et >> 5; g < G; g += 4) {
          const int64_t gg = t * G + g;
          if (gg < prm.n_groups) {
            const float* p = stage + cc * MP_STAGE_LD + g * k;
            const float b = bias_s[cb * 32 + cc];
            float res;
            if (prm.pool_mean) {
              // mean-pool (reference aggregators.py:246-273): ReLU does not commute with the mean, so bias + ReLU
              // are applied per element, summed in j order, divided by k
              float sacc = 0.f;
              for (int j = 0; j < k; ++j) sacc += fmaxf(p[j] + b, 0.f);
              res = sacc / (float)k;
            } else {
              float m = -3.0e38f;
              int j = 0;
              for (; j + 8 <= k; j += 8) {
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = p[j + u];
                m = fmaxf(m, fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7]))));
              }
              for (; j < k; ++j) m = fmaxf(m, p[j]);
              res = fmaxf(m + b, 0.f);                                          // Dense bias + ReLU (commute with the max)
            }
            prm.out[gg * prm.ldo + hcol0 + cc] = res;
          }
        }
      }
      MP_EPI_T(e4);
      named_bar_sync(1, 128);
      MP_EPI_T(e5);
      MP_EPI_ACC(0, e0, e1);
      MP_EPI_ACC(1, e1, e2);
      MP_EPI_ACC(2, e2, e3);
      MP_EPI_ACC(3, e3, e4);
      MP_EPI_ACC(4, e4, e5);
      MP_EPI_ACC(5, 0, 1);
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&acc_empty[buf]);
    if (et == 0) mp_stamp(tcount, 7);
  }
}

// MP_SA A stages of 8 KB, MP_INFLIGHT cp.async K-blocks in flight per producer thread: <7, 5> when the resident
// weights need <= 19 slots (K <= 608), else <6, 4>
// kAsyncArrive: the producers never wait for their own copies - each thread hands the stage's full barrier a
// cp.async.mbarrier.arrive.noinc, which the hardware fires when that thread's copies have landed (no cp.async group
// wait, no proxy fence, no elected arrive; the barrier counts one arrival per producer THREAD).
template <int MP_SA, int MP_INFLIGHT, bool kAsyncArrive = false>
__global__ void __launch_bounds__(MP_THREADS, 1) maxpool_mlp_kernel(const __grid_constant__ MpParams prm) {
  static_assert(MP_INFLIGHT + 2 <= MP_SA, "a stage must be free while MP_INFLIGHT copies fly and one is consumed");
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_a[MP_SA], empty_a[MP_SA], acc_full[2], acc_empty[2], b_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float bias_s[128];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* b_res = smem;                                        // resident weight slice
  unsigned char* a_ring = smem + (MP_RING - MP_SA) * MP_IMG;
  float* stage = reinterpret_cast<float*>(smem + MP_RING * MP_IMG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % prm.n_slices;
  const int64_t tile0 = blockIdx.x / prm.n_slices, tile_step = gridDim.x / prm.n_slices;
  const int kblocks = prm.kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MP_SA; ++s) {
      mbar_init(&full_a[s], kAsyncArrive ? MP_PROD_WARPS * 32 : MP_PROD_WARPS);   // one arrive per producer thread / warp
      mbar_init(&empty_a[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);        // one arrive per epilogue warp
    }
    mbar_init(&b_full, 1);
    fence_mbar_init();
  }
  if (warp == MP_PROD_WARPS) {
    tmem_alloc(&tmem_base_smem, 256);     // two 128-column fp32 accumulators
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp < MP_PROD_WARPS) {
    // =============================== gather-A producers ===============================
    const int tid = threadIdx.x;                    // 0..127
    const int c = tid & 3, r0 = tid >> 2;           // 16-byte chunk of the 64-byte row; rows r0 + 32 i, i < 4
    const int rows_valid = prm.G * prm.k;
    uint32_t it = 0;                                // running K-block counter across tiles (stage / phase)
    int pending = 0;                                // K-blocks issued but not yet published
    auto publish = [&](uint32_t which) {
      const int s = which % MP_SA;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_a[s]);
    };
    uint32_t ptile = 0;
    for (int64_t t = tile0; t < prm.n_tiles; t += tile_step, ++ptile) {
      if (threadIdx.x == 0) mp_stamp(ptile, 0);
      const __nv_bfloat16* rowp[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = r0 + 32 * i;
        const int64_t flat = t * rows_valid + r;    // index into the (group, j) row list
        rowp[i] = nullptr;
        if (r < rows_valid && flat < prm.n_groups * prm.k) {
          int64_t id = prm.row_ids ? (int64_t)prm.row_ids[flat] : prm.row0 + flat;
          if (id < 0 || id >= prm.n_rows) id = prm.n_rows - 1;
          rowp[i] = prm.table + id * prm.pitch;
        }
      }
      for (int kb = 0; kb < kblocks; ++kb, ++it) {
        const int s = it % MP_SA;
        mbar_wait(&empty_a[s], ((it / MP_SA) & 1u) ^ 1u);
        unsigned char* a_img = a_ring + (size_t)s * MP_IMG;
        const int col = kb * MP_KCOLS + c * 8;      // first bf16 column of this 16-byte piece
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int nbytes = 0;
          if (rowp[i] != nullptr && col < prm.K) nbytes = min(8, prm.K - col) * 2;
          const void* src = nbytes ? (const void*)(rowp[i] + col) : (const void*)prm.table;
          cp_async16(a_img + sw64_off(r0 + 32 * i, c), src, nbytes);
        }
        if constexpr (kAsyncArrive) {
          cp_async_mbar_arrive_noinc(&full_a[s]);
          continue;
        }
        cp_async_commit();
        ++pending;
        if (pending == MP_INFLIGHT + 1) {           // MP_INFLIGHT K-blocks stay in flight; the oldest has landed
          cp_async_wait<MP_INFLIGHT>();
          publish(it - MP_INFLIGHT);
          --pending;
        }
      }
      if (threadIdx.x == 0) mp_stamp(ptile, 1);
    }
    // drain: publish the remaining K-blocks oldest first
    while (pending > 0) {
      if (pending >= 5) cp_async_wait<4>();
      else if (pending == 4) cp_async_wait<3>();
      else if (pending == 3) cp_async_wait<2>();
      else if (pending == 2) cp_async_wait<1>();
      else cp_async_wait<0>();
      publish(it - pending);
      --pending;
    }
  } else if (warp == MP_PROD_WARPS) {
    mp_mma_role<MP_SA>(prm, lane, kblocks, tile0, tile_step, tmem_base, full_a, empty_a, acc_full, acc_empty, b_full, a_ring,
                       b_res);
  } else if (warp == MP_PROD_WARPS + 1) {
    mp_weights_role(prm, lane, slice, kblocks, b_full, b_res);
  } else {
    mp_epilogue_role(prm, (int)threadIdx.x - (MP_PROD_WARPS + 2) * 32, warp, lane, slice, tile0, tile_step, tmem_base, acc_full,
                     acc_empty, stage, bias_s);
  }
  __syncthreads();
  if (warp == MP_PROD_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant (GS_TUNING=k4_producer=1, not the default until measured): the gather-A producers use the TMA's row gather,
// `cp.async.bulk.tensor.2d.tile::gather4` - one instruction fetches a 64-byte K-block segment of FOUR table rows named
// by index and writes them, SWIZZLE_64B applied by the tensor map, straight into the UMMA stage; completion is the
// stage's mbarrier transaction count.  No per-thread cp.async group waits, no generic->async proxy fence, no per-warp
// arrive: what remains per stage is one empty-barrier wait, one expect_tx and one gather4 per lane (32 lanes x 4 rows =
// the 128-row stage).  ptxas serialises the 32 lanes of an instruction with per-lane operands (ELECT waterfall), so
// there is one producer warp PER RING SLOT, each filling whole stages, to keep that latency off the critical path.
// Producer warp w owns slot w and fills K-blocks w, w + MP_SA, ... in order: a warp that served several slots could get
// two fills ahead of a slot it last touched long ago, and an mbarrier wait only carries ONE parity bit - it would pass on
// the stale phase and overwrite a stage that has not been consumed (tools/pipeline_model.py reproduces exactly that for
// 8 warps over 7 slots; with one warp per slot every wait is at most one phase ahead).
// Columns >= K are out of bounds for the tensor map and zero-filled by the TMA; tile rows past the last fanout group
// (never read by the epilogue) fetch row 0.
// ---------------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ void tma_gather4(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int col, int r0, int r1,
                                            int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

template <int MP_SA>
__global__ void __launch_bounds__((MP_SA + 6) * 32, 1)
    maxpool_mlp_g4_kernel(const __grid_constant__ MpParams prm, const __grid_constant__ CUtensorMap tmap) {
  constexpr int MP_G4_WARPS = MP_SA;              // one producer warp per ring slot (see above)
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_a[MP_SA], empty_a[MP_SA], acc_full[2], acc_empty[2], b_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float bias_s[128];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* b_res = smem;                                        // resident weight slice
  unsigned char* a_ring = smem + (MP_RING - MP_SA) * MP_IMG;
  float* stage = reinterpret_cast<float*>(smem + MP_RING * MP_IMG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % prm.n_slices;
  const int64_t tile0 = blockIdx.x / prm.n_slices, tile_step = gridDim.x / prm.n_slices;
  const int kblocks = prm.kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MP_SA; ++s) {
      mbar_init(&full_a[s], 1);           // the expect_tx arrive of the stage's producer warp (+ 8 KB of transactions)
      mbar_init(&empty_a[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);        // one arrive per epilogue warp
    }
    mbar_init(&b_full, 1);
    fence_mbar_init();
  }
  if (warp == MP_G4_WARPS) {
    tmem_alloc(&tmem_base_smem, 256);     // two 128-column fp32 accumulators
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp < MP_G4_WARPS) {
    // =============================== gather-A producers (TMA gather4) ===============================
    const int rows_valid = prm.G * prm.k;
    const int64_t total_rows = prm.n_groups * prm.k;
    const int64_t my_tiles = tile0 < prm.n_tiles ? (prm.n_tiles - tile0 + tile_step - 1) / tile_step : 0;
    const int64_t total_it = my_tiles * kblocks;
    const int pad_row = 0;                          // tile rows past the last group: never read by the epilogue, any valid row will do
    int cur[4], nxt[4];                             // table rows of this lane's tile rows 4 lane .. 4 lane + 3
    auto load_ids = [&](int64_t tl, int (&ids)[4]) {
      const int64_t t = tile0 + tl * tile_step;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = 4 * lane + i;
        const int64_t flat = t * rows_valid + r;    // index into the (group, j) row list
        int64_t id = pad_row;
        if (tl < my_tiles && r < rows_valid && flat < total_rows) {
          id = prm.row_ids ? (int64_t)prm.row_ids[flat] : prm.row0 + flat;
          if (id < 0 || id >= prm.n_rows) id = prm.n_rows - 1;
        }
        ids[i] = (int)id;
      }
    };
    int64_t tl = 0;
    int kb = warp;
    while (kb >= kblocks) { kb -= kblocks; ++tl; }
    load_ids(tl, cur);
    load_ids(tl + 1, nxt);
    const int s = warp;                             // this warp's slot
    uint32_t fill = 0;                              // fills of the slot so far
    for (int64_t it = warp; it < total_it; it += MP_G4_WARPS, ++fill) {
      mbar_wait(&empty_a[s], (fill & 1u) ^ 1u);
      if (lane == 0) mbar_expect_tx(&full_a[s], (uint32_t)MP_IMG);
      __syncwarp();
      tma_gather4(a_ring + (size_t)s * MP_IMG + lane * 256, &tmap, &full_a[s], kb * MP_KCOLS, cur[0], cur[1], cur[2], cur[3]);
      // this warp's next K-block: MP_G4_WARPS further on, possibly in a later tile
      kb += MP_G4_WARPS;
      int adv = 0;
      while (kb >= kblocks) { kb -= kblocks; ++adv; }
      if (adv == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
        tl += 1;
        load_ids(tl + 1, nxt);
      } else if (adv > 1) {
        tl += adv;
        load_ids(tl, cur);
        load_ids(tl + 1, nxt);
      }
    }
  } else if (warp == MP_G4_WARPS) {
    mp_mma_role<MP_SA>(prm, lane, kblocks, tile0, tile_step, tmem_base, full_a, empty_a, acc_full, acc_empty, b_full, a_ring,
                       b_res);
  } else if (warp == MP_G4_WARPS + 1) {
    mp_weights_role(prm, lane, slice, kblocks, b_full, b_res);
  } else {
    mp_epilogue_role(prm, (int)threadIdx.x - (MP_G4_WARPS + 2) * 32, warp, lane, slice, tile0, tile_step, tmem_base, acc_full,
                     acc_empty, stage, bias_s);
  }
  __syncthreads();
  if (warp == MP_G4_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Variant 2 (GS_TUNING=k4_producer=2; compiled, not yet run on a GPU): as the gather4 variant, plus thread-block clusters.
// The n_slices CTAs that work on the SAME M tile (one per 128-wide slice of the hidden dimension) form a cluster of
// CL = n_slices (2 or 4) CTAs; each gathers only 128 / CL of the tile's rows and MULTICASTS them into the A stage of
// every CTA of the cluster, so an A row crosses the L2 -> SM fabric once per cluster instead of once per slice
// (that fabric's chip-wide cap, not the tensor pipe, bounds the non-multicast kernels once the proxy fence is gone).
// Hand-off: a CTA's full barrier collects its own expect_tx arrive + 8 KB of transactions from all CL issuers; its
// empty barrier has CL arrivals - every CTA's MMA warp commits with .multicast::cluster to all CL empty barriers -,
// so a producer re-fills a stage only when every CTA of the cluster has consumed it.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_gather4_multicast(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, uint16_t mask,
                                                      int col, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5, %6, %7, %8}], [%2], %3;" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "h"(mask), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

template <int MP_SA, int CL>
__global__ void __launch_bounds__((MP_SA + 6) * 32, 1)
    maxpool_mlp_g4mc_kernel(const __grid_constant__ MpParams prm, const __grid_constant__ CUtensorMap tmap) {
  static_assert(CL == 2 || CL == 4, "cluster = the 2 or 4 hidden slices of one M tile");
  constexpr int MP_G4_WARPS = MP_SA;              // one producer warp per ring slot (see the gather4 variant)
  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_a[MP_SA], empty_a[MP_SA], acc_full[2], acc_empty[2], b_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float bias_s[128];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* b_res = smem;                                        // resident weight slice
  unsigned char* a_ring = smem + (MP_RING - MP_SA) * MP_IMG;
  float* stage = reinterpret_cast<float*>(smem + MP_RING * MP_IMG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % prm.n_slices;                        // == rank in the cluster (n_slices == CL)
  const int64_t tile0 = blockIdx.x / prm.n_slices, tile_step = gridDim.x / prm.n_slices;
  const int kblocks = prm.kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MP_SA; ++s) {
      mbar_init(&full_a[s], 1);           // own expect_tx arrive; the 8 KB arrive from all CL issuers
      mbar_init(&empty_a[s], CL);         // one multicast commit per CTA of the cluster
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 4);        // one arrive per epilogue warp
    }
    mbar_init(&b_full, 1);
    fence_mbar_init();
  }
  if (warp == MP_G4_WARPS) {
    tmem_alloc(&tmem_base_smem, 256);     // two 128-column fp32 accumulators
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // every CTA's barriers exist before any remote arrive / multicast write
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp < MP_G4_WARPS) {
    // =============================== gather-A producers (TMA gather4, multicast) ===============================
    constexpr int ROWS_PER_CTA = 128 / CL;          // this CTA's share of the tile's rows
    constexpr int LANES = ROWS_PER_CTA / 4;         // one gather4 (4 rows) per active lane
    const int rank = (int)cluster_ctarank();
    const int rows_valid = prm.G * prm.k;
    const int64_t total_rows = prm.n_groups * prm.k;
    const int64_t my_tiles = tile0 < prm.n_tiles ? (prm.n_tiles - tile0 + tile_step - 1) / tile_step : 0;
    const int64_t total_it = my_tiles * kblocks;
    const int pad_row = 0;                          // tile rows past the last group: never read by the epilogue, any valid row will do
    const int row_base = rank * ROWS_PER_CTA + 4 * lane;
    int cur[4], nxt[4];
    auto load_ids = [&](int64_t tl, int (&ids)[4]) {
      const int64_t t = tile0 + tl * tile_step;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = row_base + i;
        const int64_t flat = t * rows_valid + r;
        int64_t id = pad_row;
        if (lane < LANES && tl < my_tiles && r < rows_valid && flat < total_rows) {
          id = prm.row_ids ? (int64_t)prm.row_ids[flat] : prm.row0 + flat;
          if (id < 0 || id >= prm.n_rows) id = prm.n_rows - 1;
        }
        ids[i] = (int)id;
      }
    };
    int64_t tl = 0;
    int kb = warp;
    while (kb >= kblocks) { kb -= kblocks; ++tl; }
    load_ids(tl, cur);
    load_ids(tl + 1, nxt);
    const int s = warp;                             // this warp's slot
    uint32_t fill = 0;                              // fills of the slot so far
    for (int64_t it = warp; it < total_it; it += MP_G4_WARPS, ++fill) {
      mbar_wait(&empty_a[s], (fill & 1u) ^ 1u);     // every CTA of the cluster has consumed the previous fill
      if (lane == 0) mbar_expect_tx(&full_a[s], (uint32_t)MP_IMG);
      __syncwarp();
      if (lane < LANES)
        tma_gather4_multicast(a_ring + (size_t)s * MP_IMG + (size_t)(row_base / 4) * 256, &tmap, &full_a[s],
                              (uint16_t)((1u << CL) - 1u), kb * MP_KCOLS, cur[0], cur[1], cur[2], cur[3]);
      kb += MP_G4_WARPS;
      int adv = 0;
      while (kb >= kblocks) { kb -= kblocks; ++adv; }
      if (adv == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
        tl += 1;
        load_ids(tl + 1, nxt);
      } else if (adv > 1) {
        tl += adv;
        load_ids(tl, cur);
        load_ids(tl + 1, nxt);
      }
    }
  } else if (warp == MP_G4_WARPS) {
    mp_mma_role<MP_SA, CL>(prm, lane, kblocks, tile0, tile_step, tmem_base, full_a, empty_a, acc_full, acc_empty, b_full,
                           a_ring, b_res);
  } else if (warp == MP_G4_WARPS + 1) {
    mp_weights_role(prm, lane, slice, kblocks, b_full, b_res);
  } else {
    mp_epilogue_role(prm, (int)threadIdx.x - (MP_G4_WARPS + 2) * 32, warp, lane, slice, tile0, tile_step, tmem_base, acc_full,
                     acc_empty, stage, bias_s);
  }
  __syncthreads();
  cluster_sync_all();                     // no CTA leaves while a peer may still write its stages or signal its barriers
  if (warp == MP_G4_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace gs

extern "C" {

/* developer probe (not in the public header) */
int32_t gs_debug_read_maxpool_timeline(unsigned long long* out_host, int32_t n) {
  if (n > 128) n = 128;
  GS_CUDA(cudaDeviceSynchronize());
  GS_CUDA(cudaMemcpyFromSymbol(out_host, gs::g_mp_dbg, sizeof(unsigned long long) * n));
  return GS_OK;
}

int64_t gs_maxpool_mlp_workspace_bytes(int32_t K, int32_t hidden) {
  if (K < 1 || hidden < 1) return -1;
  const int kblocks = (K + gs::MP_KCOLS - 1) / gs::MP_KCOLS, slices = (hidden + 127) / 128;
  return (int64_t)kblocks * slices * gs::MP_IMG;
}

int32_t gs_maxpool_mlp_pack(const float* Wm, int64_t ldw, int32_t K, int32_t hidden, void* workspace, void* stream) {
  GS_REQUIRE(Wm && workspace && K >= 1 && hidden >= 1 && ldw >= hidden, "gs_maxpool_mlp_pack: bad arguments");
  GS_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 127u) == 0, "gs_maxpool_mlp_pack: workspace must be 128-byte aligned");
  const int kblocks = (K + gs::MP_KCOLS - 1) / gs::MP_KCOLS, slices = (hidden + 127) / 128;
  gs::maxpool_pack_kernel<<<kblocks * slices, 256, 0, (cudaStream_t)stream>>>(Wm, ldw, K, hidden, kblocks,
                                                                             (unsigned char*)workspace);
  return gs::launch_check("maxpool_pack_kernel");
}


// bf16 table [n_rows, K] (row pitch in elements) as a 2-D tensor map for tile::gather4: box = one K-block segment
// (32 columns = 64 bytes) of ONE row - the instruction names four rows -, SWIZZLE_64B to match the UMMA stage layout
static int32_t make_table_tensor_map(CUtensorMap* out, const void* table, int64_t n_rows, int32_t K, int64_t pitch) {
  typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    GS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    GS_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available from this driver");
    encode = (encode_fn)fn;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)n_rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)pitch * 2};
  const cuuint32_t box[2] = {(cuuint32_t)gs::MP_KCOLS, 1};
  const cuuint32_t estride[2] = {1, 1};
  const CUresult rc = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(table), gdim, gstride, box, estride,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    gs::set_error("cuTensorMapEncodeTiled failed (CUresult %d) for table [%lld, %d] pitch %lld", (int)rc, (long long)n_rows, K,
                  (long long)pitch);
    return GS_ERR_CUDA;
  }
  return GS_OK;
}

static int32_t pool_mlp_fused(const void* table_bf16, int64_t n_rows, int32_t K, int64_t pitch, const int32_t* row_ids,
                              int64_t row0, int64_t n_groups, int32_t k, const void* packed_weights, const float* bias,
                              int32_t hidden, float* out, int64_t ldo, int32_t pool_mean, void* stream) {
  GS_REQUIRE(n_groups >= 0 && k >= 1, "gs_maxpool_mlp_fused: bad n_groups / k");
  if (n_groups == 0) return GS_OK;
  GS_REQUIRE(table_bf16 && packed_weights && out, "gs_maxpool_mlp_fused: NULL pointer");
  GS_REQUIRE(n_rows > 0 && n_rows < 0x7fffffffLL && K >= 1 && pitch >= K, "gs_maxpool_mlp_fused: bad table shape");
  GS_REQUIRE((pitch * 2) % 16 == 0 && (reinterpret_cast<uintptr_t>(table_bf16) & 15u) == 0,
             "gs_maxpool_mlp_fused: table rows must be 16-byte multiples and 16-byte aligned (pitch %% 8 == 0)");
  GS_REQUIRE((reinterpret_cast<uintptr_t>(packed_weights) & 127u) == 0, "gs_maxpool_mlp_fused: packed weights misaligned");
  if (k > 128 || (K + gs::MP_KCOLS - 1) / gs::MP_KCOLS > gs::MP_MAX_KB || hidden % 128 != 0) {
    gs::set_error("gs_maxpool_mlp_fused: needs k <= 128, K <= %d, hidden %% 128 == 0 (k=%d K=%d hidden=%d)",
                  gs::MP_MAX_KB * gs::MP_KCOLS, k, K, hidden);
    return GS_ERR_UNSUPPORTED;
  }
  GS_REQUIRE(ldo >= hidden, "gs_maxpool_mlp_fused: ldo < hidden");
  gs::MpParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.table = (const __nv_bfloat16*)table_bf16;
  prm.n_rows = n_rows; prm.pitch = pitch; prm.K = K; prm.kblocks = (K + gs::MP_KCOLS - 1) / gs::MP_KCOLS;
  prm.row_ids = row_ids; prm.row0 = row0; prm.n_groups = n_groups; prm.k = k; prm.G = 128 / k;
  prm.n_tiles = (n_groups + prm.G - 1) / prm.G;
  prm.hidden = hidden; prm.n_slices = hidden / 128;
  prm.wimg = (const unsigned char*)packed_weights; prm.bias = bias; prm.out = out; prm.ldo = ldo;
  prm.pool_mean = pool_mean;
  prm.issue_elect = gs::tuning("mma_issue", 1) != 0;
  {
    int32_t rc_attr = gs::ensure_dyn_smem((const void*)gs::maxpool_mlp_kernel<7, 5>, gs::MP_SMEM);
    if (rc_attr == GS_OK) rc_attr = gs::ensure_dyn_smem((const void*)gs::maxpool_mlp_kernel<6, 4>, gs::MP_SMEM);
    if (rc_attr != GS_OK) return rc_attr;
  }
  int64_t ctas = (int64_t)(gs::sm_count() / prm.n_slices) * prm.n_slices;   // a whole number of slice groups
  if (ctas < prm.n_slices) ctas = prm.n_slices;
  if (ctas > prm.n_tiles * prm.n_slices) ctas = prm.n_tiles * prm.n_slices;
  const int producer = gs::tuning("k4_producer", 0);
  if (producer == 2 && (prm.n_slices == 2 || prm.n_slices == 4) && prm.issue_elect) {
    // clusters of n_slices CTAs, TMA gather4 multicast (see maxpool_mlp_g4mc_kernel)
    CUtensorMap tmap;
    const int32_t rc = make_table_tensor_map(&tmap, table_bf16, n_rows, K, pitch);
    if (rc != GS_OK) return rc;
    const bool sa7 = prm.kblocks <= gs::MP_RING - 7;
    const void* fn = prm.n_slices == 4
                         ? (sa7 ? (const void*)gs::maxpool_mlp_g4mc_kernel<7, 4> : (const void*)gs::maxpool_mlp_g4mc_kernel<6, 4>)
                         : (sa7 ? (const void*)gs::maxpool_mlp_g4mc_kernel<7, 2> : (const void*)gs::maxpool_mlp_g4mc_kernel<6, 2>);
    const int32_t rc_attr = gs::ensure_dyn_smem(fn, gs::MP_SMEM);
    if (rc_attr != GS_OK) return rc_attr;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctas);                      // a multiple of n_slices: whole clusters
    cfg.blockDim = dim3((unsigned)(((sa7 ? 7 : 6) + 6) * 32));
    cfg.dynamicSmemBytes = gs::MP_SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = (unsigned)prm.n_slices;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    void* args[2] = {(void*)&prm, (void*)&tmap};
    GS_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    return gs::launch_check("maxpool_mlp_g4mc_kernel");
  }
  if (producer >= 1) {

    // TMA gather4 producers (see maxpool_mlp_g4_kernel)
    {
      int32_t rc_attr = gs::ensure_dyn_smem((const void*)gs::maxpool_mlp_g4_kernel<7>, gs::MP_SMEM);
      if (rc_attr == GS_OK) rc_attr = gs::ensure_dyn_smem((const void*)gs::maxpool_mlp_g4_kernel<6>, gs::MP_SMEM);
      if (rc_attr != GS_OK) return rc_attr;
    }
    CUtensorMap tmap;
    const int32_t rc = make_table_tensor_map(&tmap, table_bf16, n_rows, K, pitch);
    if (rc != GS_OK) return rc;
    if (prm.kblocks <= gs::MP_RING - 7)
      gs::maxpool_mlp_g4_kernel<7><<<(unsigned)ctas, (7 + 6) * 32, gs::MP_SMEM, (cudaStream_t)stream>>>(prm, tmap);
    else
      gs::maxpool_mlp_g4_kernel<6><<<(unsigned)ctas, (6 + 6) * 32, gs::MP_SMEM, (cudaStream_t)stream>>>(prm, tmap);
    return gs::launch_check("maxpool_mlp_g4_kernel");
  }
  if (producer == -1) {       // cp.async producers with hardware-fired arrives (no group wait / proxy fence)
    int32_t rc_attr = gs::ensure_dyn_smem((const void*)gs::maxpool_mlp_kernel<7, 5, true>, gs::MP_SMEM);
    if (rc_attr == GS_OK) rc_attr = gs::ensure_dyn_smem((const void*)gs::maxpool_mlp_kernel<6, 4, true>, gs::MP_SMEM);
    if (rc_attr != GS_OK) return rc_attr;
    if (prm.kblocks <= gs::MP_RING - 7)
      gs::maxpool_mlp_kernel<7, 5, true><<<(unsigned)ctas, gs::MP_THREADS, gs::MP_SMEM, (cudaStream_t)stream>>>(prm);
    else
      gs::maxpool_mlp_kernel<6, 4, true><<<(unsigned)ctas, gs::MP_THREADS, gs::MP_SMEM, (cudaStream_t)stream>>>(prm);
    return gs::launch_check("maxpool_mlp_kernel<async arrive>");
  }
  if (prm.kblocks <= gs::MP_RING - 7)
    gs::maxpool_mlp_kernel<7, 5><<<(unsigned)ctas, gs::MP_THREADS, gs::MP_SMEM, (cudaStream_t)stream>>>(prm);
  else
    gs::maxpool_mlp_kernel<6, 4><<<(unsigned)ctas, gs::MP_THREADS, gs::MP_SMEM, (cudaStream_t)stream>>>(prm);
  return gs::launch_check("maxpool_mlp_kernel");
}

int32_t gs_maxpool_mlp_fused(const void* table_bf16, int64_t n_rows, int32_t K, int64_t pitch, const int32_t* row_ids,
                             int64_t row0, int64_t n_groups, int32_t k, const void* packed_weights, const float* bias,
                             int32_t hidden, float* out, int64_t ldo, void* stream) {
  return pool_mlp_fused(table_bf16, n_rows, K, pitch, row_ids, row0, n_groups, k, packed_weights, bias, hidden, out, ldo, 0,
                        stream);
}

int32_t gs_meanpool_mlp_fused(const void* table_bf16, int64_t n_rows, int32_t K, int64_t pitch, const int32_t* row_ids,
                              int64_t row0, int64_t n_groups, int32_t k, const void* packed_weights, const float* bias,
                              int32_t hidden, float* out, int64_t ldo, void* stream) {
  return pool_mlp_fused(table_bf16, n_rows, K, pitch, row_ids, row0, n_groups, k, packed_weights, bias, hidden, out, ldo, 1,
                        stream);
}

}  // extern "C"

// R-MAT graph generated DIRECTLY as CSR on the device (BASELINE.json configs[4]; SURVEY section 8d config 5: a, b, c, d =
// 0.57, 0.19, 0.19, 0.05, scale 27 trimmed to 10^8 nodes, ~20 directed entries per node).  No reference counterpart: the
// reference loads its graphs from disk (graphsage/utils.py:19-75); this is the synthetic stand-in for a graph that size.
//
// R-MAT draws each edge by descending `scale` levels of a 2x2 quadrant choice with probabilities (a, b; c, d).  Factor
// that into "row first, then column given the row": the row id x has probability p(x) = (a+b)^zeros(x) (c+d)^ones(x), and
// given the row's bit at a level the column's bit there is 1 with probability b/(a+b) (row bit 0) or d/(c+d) (row bit 1).
// So a CSR can be written without sorting 2*10^9 edges:
//   degrees : deg(y) = floor(L) + [u < frac(L)],  L = edge_factor * n * (p(r) + p(r + n)),   (stochastic rounding, one draw)
//             r = the R-MAT id behind output row y; ids >= n of the 2^scale id space fold onto id - n (the trim);
//   entries : edge j of row y picks the pre-image x in {r, r + n} in proportion to p, then one 16-bit draw per level gives
//             the column's bits; column ids fold with mod n; a self loop moves to the next id.
// Output ids are scrambled by the bijection y = (x * mul + add) mod n (gcd(mul, n) = 1) because R-MAT concentrates its
// hubs on ids with few set bits - contiguous-range partitions would put every hub on GPU 0.
// RNG contract (oracle/rmat.py): Philox4x32-10, key = seed;
//   degree draw of row y : word 0 of block (ctr = (y, 0, 0, TAG_DEG));
//   edge j of row y      : half-word h (0..31) = 16 bits of word (h >> 1) & 3 of block (ctr = (j, h >> 3, y, TAG_FILL)),
//                          low half first; h = 0 picks the pre-image, h = 1 + level the column bit of that level.
#include "common.cuh"

namespace gs {

constexpr uint32_t kRmatTagDeg = 0x08000000u;
constexpr uint32_t kRmatTagFill = 0x08000001u;

struct RmatParams {
  int32_t scale;
  int64_t n;
  double lambda_scale;          // edge_factor * n
  double prow[33];              // prow[z] = (a+b)^(scale - z) * (c+d)^z : probability of a row id with z set bits
  uint32_t thr0, thr1;          // 65536 * P(column bit = 1 | row bit = 0 / 1), truncated
  uint64_t seed, mul, mul_inv, add;
};

__host__ __device__ __forceinline__ int64_t rmat_unscramble(const RmatParams& p, int64_t y) {
  const uint64_t t = (uint64_t)y >= p.add ? (uint64_t)y - p.add : (uint64_t)y + (uint64_t)p.n - p.add;
  return (int64_t)((t * p.mul_inv) % (uint64_t)p.n);
}
__host__ __device__ __forceinline__ int64_t rmat_scramble(const RmatParams& p, int64_t x) {
  return (int64_t)(((uint64_t)x * p.mul + p.add) % (uint64_t)p.n);
}

__global__ void __launch_bounds__(256) rmat_degrees_kernel(const __grid_constant__ RmatParams p, int32_t* __restrict__ deg) {
  for (int64_t y = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; y < p.n; y += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = rmat_unscramble(p, y);
    double pr = p.prow[__popcll((unsigned long long)r)];
    const int64_t r2 = r + p.n;
    if (r2 < ((int64_t)1 << p.scale)) pr = pr + p.prow[__popcll((unsigned long long)r2)];
    const double lam = p.lambda_scale * pr;
    const double fl = floor(lam);
    u32x4 ctr{(uint32_t)y, 0u, 0u, kRmatTagDeg};
    const u32x4 rnd = philox4x32_10(ctr, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    const double u = ((double)rnd.x + 0.5) * (1.0 / 4294967296.0);
    double d = fl + ((u < (lam - fl)) ? 1.0 : 0.0);
    if (d > 2147483647.0) d = 2147483647.0;
    deg[y] = (int32_t)d;
  }
}

// entry j of row y (start = indptr[y]); r / r2 / pick_thr describe the row's pre-images
__device__ __forceinline__ void rmat_entry(const RmatParams& p, int64_t y, int64_t r, int64_t r2, uint32_t pick_thr, int64_t j,
                                           int64_t start, int32_t* __restrict__ indices) {
  uint32_t hw[32];
#pragma unroll
  for (int blk = 0; blk < 4; ++blk) {
    u32x4 ctr{(uint32_t)j, (uint32_t)blk, (uint32_t)y, kRmatTagFill};
    const u32x4 rnd = philox4x32_10(ctr, (uint32_t)p.seed, (uint32_t)(p.seed >> 32));
    const uint32_t w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      hw[blk * 8 + 2 * q] = w[q] & 0xffffu;
      hw[blk * 8 + 2 * q + 1] = w[q] >> 16;
    }
  }
  const int64_t x = (hw[0] < pick_thr) ? r2 : r;
  int64_t col = 0;
#pragma unroll
  for (int l = 0; l < 31; ++l) {
    if (l < p.scale) {
      const int rb = (int)((x >> (p.scale - 1 - l)) & 1);
      const int cb = hw[1 + l] < (rb ? p.thr1 : p.thr0) ? 1 : 0;
      col = (col << 1) | cb;
    }
  }
  col %= p.n;
  int64_t cy = rmat_scramble(p, col);
  if (cy == y) cy = (cy + 1) % p.n;
  indices[start + j] = (int32_t)cy;
}

__device__ __forceinline__ uint32_t rmat_pick_thr(const RmatParams& p, int64_t r, int64_t r2) {
  if (r2 >= ((int64_t)1 << p.scale)) return 0u;                      // P(pre-image = r + n), as a 16-bit threshold
  const double p1 = p.prow[__popcll((unsigned long long)r)], p2 = p.prow[__popcll((unsigned long long)r2)];
  return (uint32_t)(65536.0 * (p2 / (p1 + p2)));
}

// one warp per row, lanes stride over the row's entries; rows longer than `long_threshold` are left to rmat_fill_long_kernel
// (R-MAT's hubs: at scale 27 one row has 1.2 M entries - a single warp would need half a minute for it)
__global__ void __launch_bounds__(256) rmat_fill_kernel(const __grid_constant__ RmatParams p, const int64_t* __restrict__ indptr,
                                                        int32_t* __restrict__ indices, int64_t long_threshold) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t y = warp; y < p.n; y += nwarps) {
    const int64_t start = indptr[y], deg = indptr[y + 1] - start;
    if (deg <= 0 || deg > long_threshold) continue;
    const int64_t r = rmat_unscramble(p, y), r2 = r + p.n;
    const uint32_t pick_thr = rmat_pick_thr(p, r, r2);
    for (int64_t j = lane; j < deg; j += 32) rmat_entry(p, y, r, r2, pick_thr, j, start, indices);
  }
}

// the long rows (ids listed by the caller): blockIdx.y = which long row, the x-dimension strides over its entries
__global__ void __launch_bounds__(256) rmat_fill_long_kernel(const __grid_constant__ RmatParams p, const int64_t* __restrict__ indptr,
                                                             const int64_t* __restrict__ long_rows, int32_t* __restrict__ indices) {
  const int64_t y = long_rows[blockIdx.y];
  const int64_t start = indptr[y], deg = indptr[y + 1] - start;
  const int64_t r = rmat_unscramble(p, y), r2 = r + p.n;
  const uint32_t pick_thr = rmat_pick_thr(p, r, r2);
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < deg; j += (int64_t)gridDim.x * blockDim.x)
    rmat_entry(p, y, r, r2, pick_thr, j, start, indices);
}

static int32_t fill_params(RmatParams& p, int32_t scale, int64_t n, double edge_factor, double a, double b, double c, double d,
                           uint64_t seed, uint64_t mul, uint64_t mul_inv, uint64_t add) {
  GS_REQUIRE(scale >= 1 && scale <= 31, "gs_rmat: scale must be in [1, 31] (got %d)", scale);
  GS_REQUIRE(n >= 2 && n <= ((int64_t)1 << scale) && n < 0x7fffffffLL, "gs_rmat: n_nodes must be in [2, 2^scale]");
  GS_REQUIRE(a > 0 && b > 0 && c > 0 && d > 0 && fabs(a + b + c + d - 1.0) < 1e-9, "gs_rmat: a + b + c + d must be 1");
  GS_REQUIRE(mul > 0 && mul < (uint64_t)n && add < (uint64_t)n && (mul * mul_inv) % (uint64_t)n == 1,
             "gs_rmat: (mul, mul_inv, add) is not a bijection of [0, n)");
  memset(&p, 0, sizeof(p));
  p.scale = scale; p.n = n; p.lambda_scale = edge_factor * (double)n;
  for (int z = 0; z <= scale; ++z) {           // plain repeated multiplication: the oracle does the same, bit for bit
    double v = 1.0;
    for (int i = 0; i < scale - z; ++i) v = v * (a + b);
    for (int i = 0; i < z; ++i) v = v * (c + d);
    p.prow[z] = v;
  }
  p.thr0 = (uint32_t)(65536.0 * (b / (a + b)));
  p.thr1 = (uint32_t)(65536.0 * (d / (c + d)));
  p.seed = seed; p.mul = mul; p.mul_inv = mul_inv; p.add = add;
  return GS_OK;
}

}  // namespace gs

extern "C" {

int32_t gs_rmat_degrees(int32_t scale, int64_t n_nodes, double edge_factor, double a, double b, double c, double d,
                        uint64_t seed, uint64_t mul, uint64_t mul_inv, uint64_t add, int32_t* deg_out, void* stream) {
  gs::RmatParams p;
  const int32_t rc = gs::fill_params(p, scale, n_nodes, edge_factor, a, b, c, d, seed, mul, mul_inv, add);
  if (rc != GS_OK) return rc;
  GS_REQUIRE(deg_out != nullptr && edge_factor > 0, "gs_rmat_degrees: bad arguments");
  int64_t blocks = (n_nodes + 255) / 256;
  int64_t cap = (int64_t)gs::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  gs::rmat_degrees_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, deg_out);
  return gs::launch_check("rmat_degrees_kernel");
}

int32_t gs_rmat_fill(int32_t scale, int64_t n_nodes, double a, double b, double c, double d, uint64_t seed, uint64_t mul,
                     uint64_t mul_inv, uint64_t add, const int64_t* indptr, int32_t* indices, const int64_t* long_rows,
                     int64_t n_long, int64_t long_threshold, void* stream) {
  gs::RmatParams p;
  const int32_t rc = gs::fill_params(p, scale, n_nodes, 1.0, a, b, c, d, seed, mul, mul_inv, add);
  if (rc != GS_OK) return rc;
  GS_REQUIRE(indptr != nullptr && indices != nullptr, "gs_rmat_fill: NULL pointer");
  GS_REQUIRE(n_long >= 0 && n_long <= 65535 && (n_long == 0 || (long_rows != nullptr && long_threshold > 0)),
             "gs_rmat_fill: bad long-row list (at most 65535 rows; raise the threshold)");
  int64_t blocks = (n_nodes + 7) / 8;
  int64_t cap = (int64_t)gs::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  gs::rmat_fill_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, indptr, indices,
                                                                          n_long > 0 ? long_threshold : (int64_t)1 << 62);
  int32_t rc2 = gs::launch_check("rmat_fill_kernel");
  if (rc2 != GS_OK || n_long == 0) return rc2;
  gs::rmat_fill_long_kernel<<<dim3(64, (unsigned)n_long), 256, 0, (cudaStream_t)stream>>>(p, indptr, long_rows, indices);
  return gs::launch_check("rmat_fill_long_kernel");
}

}  // extern "C"

// Stable LSD radix sort of (key, index) pairs + device-wide exclusive scan.
//
// This is the "sort" half of the sort-and-segment gradient reduce: item rows are
// grouped by user id / recipe id / health-label id so that every unique table row
// is reduced by one warp in batch order -- the same order TF's
// unsorted_segment_sum uses on CPU (optimizer.py:_deduplicate_indexed_slices).
// Stability is what preserves batch order inside a segment.
//
// Per 8-bit pass: (1) per-tile digit histogram, (2) exclusive scan of the
// bin-major [256 x ntiles] table, (3) scatter with warp-level stable ranking
// (__match_any_sync multisplit).  All loads are coalesced; the tile is re-read by
// the scatter pass (L2 hit: a tile is 16 KB).
#include "common.cuh"
#include "internal.h"

namespace fr {

unsigned long long g_launches = 0;

__device__ __forceinline__ uint32_t resolve_n(const uint32_t* n_dev, uint32_t n_host) {
  if (n_dev) { const uint32_t v = *n_dev; return v < n_host ? v : n_host; }
  return n_host;
}

__global__ void __launch_bounds__(FR_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n_host, const uint32_t* n_dev,
                  int shift, uint32_t ntiles, uint32_t* __restrict__ tile_hist) {
  __shared__ uint32_t hist[RADIX_BINS];
  const uint32_t n = resolve_n(n_dev, n_host);
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = tile * SORT_TILE;
#pragma unroll
    for (int r = 0; r < SORT_TILE / FR_THREADS; ++r) {
      const uint32_t idx = base + r * FR_THREADS + threadIdx.x;
      if (idx < n) atomicAdd(&hist[(keys[idx] >> shift) & (RADIX_BINS - 1)], 1u);
    }
    __syncthreads();
    tile_hist[threadIdx.x * ntiles + tile] = hist[threadIdx.x];
    __syncthreads();
  }
}

// Each warp owns 256 consecutive keys of the tile (8 rounds of 32), so the order
// (warp, round, lane) is the input order.
__global__ void __launch_bounds__(FR_THREADS)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                     uint32_t n_host, const uint32_t* n_dev, int shift, uint32_t ntiles,
                     const uint32_t* __restrict__ tile_off, int fused_scan) {
  __shared__ uint32_t wcnt[FR_WARPS_PER_BLOCK][RADIX_BINS];
  __shared__ uint32_t wtot[FR_WARPS_PER_BLOCK];
  const uint32_t n = resolve_n(n_dev, n_host);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  constexpr int ROUNDS = SORT_TILE / FR_THREADS;  // 8
  for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint32_t base = tile * SORT_TILE + warp * (ROUNDS * 32);
    if (tile * SORT_TILE >= n) break;   // uniform: later tiles are empty too
    for (int b = lane; b < RADIX_BINS; b += 32) wcnt[warp][b] = 0;
    __syncwarp();
    uint32_t key[ROUNDS], val[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const uint32_t idx = base + r * 32 + lane;
      const bool valid = idx < n;
      key[r] = valid ? keys_in[idx] : 0u;
      val[r] = valid ? (vals_in ? vals_in[idx] : idx) : 0u;
      const uint32_t d = valid ? ((key[r] >> shift) & (RADIX_BINS - 1)) : (RADIX_BINS + lane);
      const uint32_t mask = __match_any_sync(FR_FULL, d);
      if (valid && (mask & lt) == 0) wcnt[warp][d] += __popc(mask);   // group leader
      __syncwarp();
    }
    __syncthreads();
    {  // exclusive scan over warps per digit, seeded with the global tile offset
      const int b = threadIdx.x;
      uint32_t run;
      if (fused_scan) {
        // few tiles: tile_off holds the RAW per-tile histograms and every block derives
        // its own offsets (saves the separate scan launches of the pass):
        //   offset[b][tile] = sum_{b'<b} total[b'] + sum_{t<tile} hist[b][t]
        const uint32_t* hrow = tile_off + (size_t)b * ntiles;
        uint32_t pre = 0, tot = 0;
        for (uint32_t t = 0; t < ntiles; ++t) {
          const uint32_t c = hrow[t];
          tot += c;
          if (t < tile) pre += c;
        }
        uint32_t inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t y = __shfl_up_sync(FR_FULL, inc, o);
          if (lane >= o) inc += y;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0;
#pragma unroll
        for (int w = 0; w < FR_WARPS_PER_BLOCK; ++w) if (w < warp) wbase += wtot[w];
        run = wbase + (inc - tot) + pre;
      } else {
        run = tile_off[b * ntiles + tile];
      }
#pragma unroll
      for (int w = 0; w < FR_WARPS_PER_BLOCK; ++w) {
        const uint32_t c = wcnt[w][b];
        wcnt[w][b] = run;
        run += c;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const uint32_t idx = base + r * 32 + lane;
      const bool valid = idx < n;
      const uint32_t d = valid ? ((key[r] >> shift) & (RADIX_BINS - 1)) : (RADIX_BINS + lane);
      const uint32_t mask = __match_any_sync(FR_FULL, d);
      uint32_t dst = 0;
      if (valid) dst = wcnt[warp][d] + __popc(mask & lt);
      __syncwarp();
      if (valid && (mask & lt) == 0) wcnt[warp][d] += __popc(mask);
      __syncwarp();
      if (valid) { keys_out[dst] = key[r]; vals_out[dst] = val[r]; }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------- scan
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

// exclusive scan of SCAN_TILE values held 4 per thread (blocked); returns tile total.
__device__ __forceinline__ uint32_t block_scan_tile(uint32_t (&v)[SCAN_ITEMS], uint32_t carry) {
  __shared__ uint32_t wsum[32];
  __shared__ uint32_t total_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { const uint32_t x = v[i]; v[i] = t; t += x; }
  uint32_t inc = t;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(FR_FULL, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = wsum[lane];
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(FR_FULL, winc, o);
      if (lane >= o) winc += y;
    }
    wsum[lane] = winc - w;
    if (lane == 31) total_s = winc;
  }
  __syncthreads();
  const uint32_t off = carry + wsum[warp] + (inc - t);
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) v[i] += off;
  const uint32_t total = total_s;
  __syncthreads();
  return total;
}

__device__ __forceinline__ void scan_load(const uint32_t* in, uint32_t base, uint32_t n, uint32_t (&v)[SCAN_ITEMS]) {
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const uint32_t idx = base + threadIdx.x * SCAN_ITEMS + i;
    v[i] = idx < n ? in[idx] : 0u;
  }
}
__device__ __forceinline__ void scan_store(uint32_t* out, uint32_t base, uint32_t n, const uint32_t (&v)[SCAN_ITEMS]) {
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    const uint32_t idx = base + threadIdx.x * SCAN_ITEMS + i;
    if (idx < n) out[idx] = v[i];
  }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_single_kernel(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* total_out) {
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n; base += SCAN_TILE) {
    uint32_t v[SCAN_ITEMS];
    scan_load(in, base, n, v);
    const uint32_t tot = block_scan_tile(v, carry);
    scan_store(out, base, n, v);
    carry += tot;
  }
  if (total_out && threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_sums_kernel(const uint32_t* in, uint32_t n, uint32_t* sums) {
  uint32_t v[SCAN_ITEMS];
  scan_load(in, blockIdx.x * SCAN_TILE, n, v);
  const uint32_t tot = block_scan_tile(v, 0);
  if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const uint32_t* in, uint32_t* out, uint32_t n, const uint32_t* sums_scanned) {
  uint32_t v[SCAN_ITEMS];
  scan_load(in, blockIdx.x * SCAN_TILE, n, v);
  block_scan_tile(v, sums_scanned[blockIdx.x]);
  scan_store(out, blockIdx.x * SCAN_TILE, n, v);
}

void exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp,
                        uint32_t* total_out, cudaStream_t st) {
  if (n == 0) {
    if (total_out) cudaMemsetAsync(total_out, 0, sizeof(uint32_t), st);
    return;
  }
  if (n <= 4 * SCAN_TILE) {
    scan_single_kernel<<<1, SCAN_THREADS, 0, st>>>(in, out, n, total_out);
    g_launches += 1;
    return;
  }
  const uint32_t nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  scan_tile_sums_kernel<<<nb, SCAN_THREADS, 0, st>>>(in, n, tmp);
  scan_single_kernel<<<1, SCAN_THREADS, 0, st>>>(tmp, tmp, nb, total_out);
  scan_apply_kernel<<<nb, SCAN_THREADS, 0, st>>>(in, out, n, tmp);
  g_launches += 3;
}

int radix_sort_pairs(SortBufs& bufs, const uint32_t* keys_in, uint32_t n_host,
                     const uint32_t* n_dev, int nbits, cudaStream_t st, int sm_count) {
  int passes = (nbits + RADIX_BITS - 1) / RADIX_BITS;
  if (passes < 1) passes = 1;
  const uint32_t ntiles = (n_host + SORT_TILE - 1) / SORT_TILE;
  if (ntiles == 0) return 0;
  const int grid = (int)(ntiles < (uint32_t)(sm_count * 4) ? ntiles : (uint32_t)(sm_count * 4));
  const uint32_t* kin = keys_in;
  const uint32_t* vin = nullptr;
  int dst = 0;
  for (int p = 0; p < passes; ++p) {
    const int shift = p * RADIX_BITS;
    const int fused = ntiles <= 128 ? 1 : 0;
    radix_hist_kernel<<<grid, FR_THREADS, 0, st>>>(kin, n_host, n_dev, shift, ntiles, bufs.tile_hist);
    if (!fused) exclusive_scan_u32(bufs.tile_hist, bufs.tile_hist, RADIX_BINS * ntiles, bufs.scan_tmp, nullptr, st);
    radix_scatter_kernel<<<grid, FR_THREADS, 0, st>>>(kin, vin, bufs.k[dst], bufs.v[dst], n_host, n_dev,
                                                      shift, ntiles, bufs.tile_hist, fused);
    g_launches += 2;
    kin = bufs.k[dst];
    vin = bufs.v[dst];
    dst ^= 1;
  }
  return dst ^ 1;
}

}  // namespace fr

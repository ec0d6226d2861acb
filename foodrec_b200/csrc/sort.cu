// Stable LSD radix sort of (key, index) pairs + device-wide exclusive scan.
//
// This is the "sort" half of the sort-and-segment gradient reduce: item rows are
// grouped by user id / recipe id / health-label id so that every unique table row
// is reduced by one warp in batch order -- the same order TF's
// unsorted_segment_sum uses on CPU (optimizer.py:_deduplicate_indexed_slices).
// Stability is what preserves batch order inside a segment.
//
// Per pass (digit of up to 10 bits: 20-bit user ids sort in 2 passes): (1) per-tile digit
// histogram, (2) offsets -- derived inside the scatter kernel from the raw histograms when
// there are few tiles, else by a separate scan of the bin-major [bins x ntiles] table,
// (3) scatter with warp-level stable ranking (__match_any_sync multisplit).  One launch
// serves up to two independent jobs (blockIdx.y): the by-user and the by-recipe sort of a
// step share their launches.  All loads are coalesced; the tile is re-read by the scatter
// pass (L2 hit: a tile is 16 KB).
#include "common.cuh"
#include "internal.h"

namespace fr {

unsigned long long g_launches = 0;

__device__ __forceinline__ uint32_t resolve_n(const uint32_t* n_dev, uint32_t n_host) {
  if (n_dev) { const uint32_t v = *n_dev; return v < n_host ? v : n_host; }
  return n_host;
}

constexpr int MAX_DIGIT_BITS = 10;
constexpr int MAX_BINS = 1 << MAX_DIGIT_BITS;
constexpr int SCAN_TILE_C = 4096;          // = SCAN_TILE below (entries of the offset table one scan block covers)
constexpr int FUSED_SCAN_MAX_TILES = 16;   // every block re-reads bins x tiles counters: measured 0.30 ms per step at 128 tiles vs 0.145 ms with the separate scan at 256

struct PassJob {
  const uint32_t* keys_in; const uint32_t* vals_in;   // vals_in == nullptr: identity
  uint32_t* keys_out; uint32_t* vals_out;
  uint32_t n_host; const uint32_t* n_dev;
  int shift, bits;                                     // bits == 0: job idle in this pass
  uint32_t ntiles; uint32_t* tile_hist; int fused;
  const uint32_t* blk_pref;                            // separate scan: exclusive prefix of the scan blocks (see scan_offsets_kernel)
};
struct PassJobs { PassJob j[2]; };

__global__ void __launch_bounds__(FR_THREADS)
radix_hist_kernel(const PassJobs jobs) {
  __shared__ uint32_t hist[MAX_BINS];
  const PassJob& J = jobs.j[blockIdx.y];
  if (J.bits == 0) return;
  const uint32_t n = resolve_n(J.n_dev, J.n_host);
  const uint32_t bins = 1u << J.bits, mask = bins - 1u;
  // fused offsets only ever read the tiles that hold keys; the separate scan reads them all
  const uint32_t nt = J.fused ? min(J.ntiles, (n + SORT_TILE - 1) / SORT_TILE) : J.ntiles;
  for (uint32_t tile = blockIdx.x; tile < nt; tile += gridDim.x) {
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    const uint32_t base = tile * SORT_TILE;
#pragma unroll
    for (int r = 0; r < SORT_TILE / FR_THREADS; ++r) {
      const uint32_t idx = base + r * FR_THREADS + threadIdx.x;
      if (idx < n) atomicAdd(&hist[(J.keys_in[idx] >> J.shift) & mask], 1u);
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < bins; b += blockDim.x) J.tile_hist[(size_t)b * J.ntiles + tile] = hist[b];
    __syncthreads();
  }
}

// Each warp owns 256 consecutive keys of the tile (8 rounds of 32), so the order
// (warp, round, lane) is the input order.
__global__ void __launch_bounds__(FR_THREADS)
radix_scatter_kernel(const PassJobs jobs) {
  __shared__ uint32_t wcnt[FR_WARPS_PER_BLOCK][MAX_BINS];
  __shared__ uint32_t wtot[FR_WARPS_PER_BLOCK];
  const PassJob& J = jobs.j[blockIdx.y];
  if (J.bits == 0) return;
  const uint32_t n = resolve_n(J.n_dev, J.n_host);
  const uint32_t bins = 1u << J.bits, mask = bins - 1u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  constexpr int ROUNDS = SORT_TILE / FR_THREADS;  // 8
  constexpr int BPT = MAX_BINS / FR_THREADS;      // bins per thread (contiguous) in the offset scan
  for (uint32_t tile = blockIdx.x; tile < J.ntiles; tile += gridDim.x) {
    const uint32_t base = tile * SORT_TILE + warp * (ROUNDS * 32);
    if (tile * SORT_TILE >= n) break;   // uniform: later tiles are empty too
    for (uint32_t b = lane; b < bins; b += 32) wcnt[warp][b] = 0;
    __syncwarp();
    uint32_t key[ROUNDS], val[ROUNDS];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const uint32_t idx = base + r * 32 + lane;
      const bool valid = idx < n;
      key[r] = valid ? J.keys_in[idx] : 0u;
      val[r] = valid ? (J.vals_in ? J.vals_in[idx] : idx) : 0u;
      const uint32_t d = valid ? ((key[r] >> J.shift) & mask) : (MAX_BINS + lane);
      const uint32_t m = __match_any_sync(FR_FULL, d);
      if (valid && (m & lt) == 0) wcnt[warp][d] += __popc(m);   // group leader
      __syncwarp();
    }
    __syncthreads();
    {  // per digit: exclusive scan over warps, seeded with the global offset of (digit, tile)
      uint32_t run[BPT];
      const uint32_t b0 = threadIdx.x * BPT;
      if (J.fused) {
        // few tiles: tile_hist holds the RAW per-tile histograms and every block derives its
        // own offsets:  offset[b][tile] = sum_{b'<b} total[b'] + sum_{t<tile} hist[b][t]
        uint32_t pre[BPT], tot[BPT], tsum = 0;
#pragma unroll
        for (int q = 0; q < BPT; ++q) {
          pre[q] = 0; tot[q] = 0;
          const uint32_t b = b0 + q;
          if (b < bins) {
            const uint32_t* hrow = J.tile_hist + (size_t)b * J.ntiles;
            const uint32_t nt_act = min(J.ntiles, (n + SORT_TILE - 1) / SORT_TILE);
            for (uint32_t t = 0; t < nt_act; ++t) {
              const uint32_t c = hrow[t];
              tot[q] += c;
              if (t < tile) pre[q] += c;
            }
          }
          tsum += tot[q];
        }
        uint32_t inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t y = __shfl_up_sync(FR_FULL, inc, o);
          if (lane >= o) inc += y;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        uint32_t acc = inc - tsum;
#pragma unroll
        for (int w = 0; w < FR_WARPS_PER_BLOCK; ++w) if (w < warp) acc += wtot[w];
#pragma unroll
        for (int q = 0; q < BPT; ++q) { run[q] = acc + pre[q]; acc += tot[q]; }
      } else {
#pragma unroll
        for (int q = 0; q < BPT; ++q) {
          const size_t e = (size_t)(b0 + q) * J.ntiles + tile;      // block-local scan + the prefix of its scan block
          run[q] = (b0 + q < bins) ? J.tile_hist[e] + J.blk_pref[e / SCAN_TILE_C] : 0u;
        }
      }
#pragma unroll
      for (int q = 0; q < BPT; ++q) {
        const uint32_t b = b0 + q;
        if (b < bins) {
          uint32_t r2 = run[q];
#pragma unroll
          for (int w = 0; w < FR_WARPS_PER_BLOCK; ++w) {
            const uint32_t c = wcnt[w][b];
            wcnt[w][b] = r2;
            r2 += c;
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      const uint32_t idx = base + r * 32 + lane;
      const bool valid = idx < n;
      const uint32_t d = valid ? ((key[r] >> J.shift) & mask) : (MAX_BINS + lane);
      const uint32_t m = __match_any_sync(FR_FULL, d);
      uint32_t dst = 0;
      if (valid) dst = wcnt[warp][d] + __popc(m & lt);
      __syncwarp();
      if (valid && (m & lt) == 0) wcnt[warp][d] += __popc(m);
      __syncwarp();
      if (valid) { J.keys_out[dst] = key[r]; J.vals_out[dst] = val[r]; }
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

// The offset scan of a radix pass in ONE launch, for up to two jobs (blockIdx.y).  Every block scans its SCAN_TILE
// entries of the bin-major [bins x ntiles] histogram in place and leaves its total; the block that finishes LAST (a
// ticket counter) turns the totals into exclusive prefixes.  The table is NOT rewritten with the prefixes added: the
// scatter kernel adds blk_pref[entry / SCAN_TILE] when it reads an entry.  (Three launches before -- tile sums, a
// one-block scan of them, apply -- on the critical path of every pass.)
static_assert(SCAN_TILE == SCAN_TILE_C, "scatter kernel's block size of the offset scan");
struct Scan2 { uint32_t* data[2]; uint32_t n[2]; uint32_t* tmp[2]; uint32_t* ticket[2]; };
__global__ void __launch_bounds__(SCAN_THREADS)
scan_offsets_kernel(const Scan2 s) {
  __shared__ bool last;
  const int j = blockIdx.y;
  const uint32_t n = s.n[j], nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (blockIdx.x >= nb) return;
  {
    uint32_t v[SCAN_ITEMS];
    scan_load(s.data[j], blockIdx.x * SCAN_TILE, n, v);
    const uint32_t tot = block_scan_tile(v, 0);
    scan_store(s.data[j], blockIdx.x * SCAN_TILE, n, v);
    if (threadIdx.x == 0) {
      s.tmp[j][blockIdx.x] = tot;
      __threadfence();
      last = atomicAdd(s.ticket[j], 1u) == nb - 1;
    }
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nb; base += SCAN_TILE) {
    uint32_t v[SCAN_ITEMS];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
      const uint32_t idx = base + threadIdx.x * SCAN_ITEMS + i;
      v[i] = idx < nb ? __ldcg(s.tmp[j] + idx) : 0u;
    }
    const uint32_t tot = block_scan_tile(v, carry);
    scan_store(s.tmp[j], base, nb, v);
    carry += tot;
  }
  if (threadIdx.x == 0) *s.ticket[j] = 0u;         // ready for the next launch
}
static void scan_offsets(const Scan2& s, int njobs, cudaStream_t st) {
  const uint32_t nmax = s.n[0] > s.n[1] ? s.n[0] : s.n[1];
  const uint32_t nb = (nmax + SCAN_TILE - 1) / SCAN_TILE;
  scan_offsets_kernel<<<dim3(nb, njobs), SCAN_THREADS, 0, st>>>(s);
  g_launches += 1;
}

// ---------------------------------------------------------------- host side
static void plan_digits(int nbits, int& passes, int& bits_per_pass) {
  if (nbits < 1) nbits = 1;
  passes = (nbits + MAX_DIGIT_BITS - 1) / MAX_DIGIT_BITS;
  bits_per_pass = (nbits + passes - 1) / passes;
}

// Sort up to two independent jobs with shared launches.  result[i] = index of the buffer
// pair (bufs.k[r], bufs.v[r]) that holds job i's sorted output.
void radix_sort_jobs(SortJob* jobs, int njobs, cudaStream_t st, int sm_count) {
  int passes[2] = {0, 0}, bpp[2] = {0, 0}, dst[2] = {0, 0};
  const uint32_t* kin[2]; const uint32_t* vin[2];
  int maxp = 0;
  uint32_t max_tiles = 0;
  for (int i = 0; i < njobs; ++i) {
    plan_digits(jobs[i].nbits, passes[i], bpp[i]);
    if (jobs[i].n_host == 0) passes[i] = 0;
    if (passes[i] > maxp) maxp = passes[i];
    kin[i] = jobs[i].keys_in; vin[i] = nullptr;
    const uint32_t nt = (jobs[i].n_host + SORT_TILE - 1) / SORT_TILE;
    if (nt > max_tiles) max_tiles = nt;
    jobs[i].result = 0;
  }
  if (maxp == 0 || max_tiles == 0) return;
  const int gx = (int)(max_tiles < (uint32_t)(sm_count * 4) ? max_tiles : (uint32_t)(sm_count * 4));
  for (int p = 0; p < maxp; ++p) {
    PassJobs pj{};
    for (int i = 0; i < 2; ++i) {
      PassJob& J = pj.j[i];
      if (i >= njobs || p >= passes[i]) { J.bits = 0; continue; }
      SortBufs& b = *jobs[i].bufs;
      J.keys_in = kin[i]; J.vals_in = vin[i]; J.keys_out = b.k[dst[i]]; J.vals_out = b.v[dst[i]];
      J.n_host = jobs[i].n_host; J.n_dev = jobs[i].n_dev;
      J.shift = p * bpp[i];
      const int left = jobs[i].nbits - J.shift;
      J.bits = left < bpp[i] ? (left < 1 ? 1 : left) : bpp[i];
      J.ntiles = (jobs[i].n_host + SORT_TILE - 1) / SORT_TILE;
      J.tile_hist = b.tile_hist;
      J.fused = J.ntiles <= FUSED_SCAN_MAX_TILES ? 1 : 0;
    }
    radix_hist_kernel<<<dim3(gx, njobs), FR_THREADS, 0, st>>>(pj);
    ++g_launches;
    Scan2 sc{};
    bool any_scan = false;
    for (int i = 0; i < njobs; ++i) {
      if (!pj.j[i].bits || pj.j[i].fused) continue;        // (n = 0: the job's blocks exit at once)
      sc.data[i] = pj.j[i].tile_hist; sc.n[i] = (uint32_t)(1u << pj.j[i].bits) * pj.j[i].ntiles;
      sc.tmp[i] = jobs[i].bufs->scan_tmp; sc.ticket[i] = jobs[i].bufs->ticket;
      pj.j[i].blk_pref = jobs[i].bufs->scan_tmp;
      any_scan = true;
    }
    if (any_scan) scan_offsets(sc, njobs, st);
    radix_scatter_kernel<<<dim3(gx, njobs), FR_THREADS, 0, st>>>(pj);
    ++g_launches;
    for (int i = 0; i < njobs; ++i) {
      if (!pj.j[i].bits) continue;
      kin[i] = pj.j[i].keys_out; vin[i] = pj.j[i].vals_out;
      jobs[i].result = dst[i];
      dst[i] ^= 1;
    }
  }
}

int radix_sort_pairs(SortBufs& bufs, const uint32_t* keys_in, uint32_t n_host,
                     const uint32_t* n_dev, int nbits, cudaStream_t st, int sm_count) {
  SortJob j{&bufs, keys_in, n_host, n_dev, nbits, 0};
  radix_sort_jobs(&j, 1, st, sm_count);
  return j.result;
}

}  // namespace fr

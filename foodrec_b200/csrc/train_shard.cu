// Helper kernels of the row-sharded training step (one process per GPU):
//   P is sharded by user % W (samples are routed to the user owner when they are loaded),
//   R by recipe % W.  A rank requests the recipe rows its batch needs, the owners answer
//   (all-to-all over NVLink), and finished gradient rows travel the other way.
// Everything here is index bookkeeping; the row movers are the kernels of train_fwd.cu /
// train_seg.cu, pointed at the receive buffer (R := rbuf, items := slot ids).
#include "common.cuh"
#include "train.cuh"

namespace fr {

// owner-major key: sorting rows by it groups the unique recipes by owner, contiguously
__global__ void shard_prep_kernel(const ShardPlanParams p) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < p.S; r += gridDim.x * blockDim.x) {
    const uint32_t it = (uint32_t)p.items[r];
    p.okeys[r] = (it % (uint32_t)p.W) * p.items_per_rank + it / (uint32_t)p.W;
    p.cats_row[r] = p.cats_in ? p.cats_in[r] : __ldg(p.item_cats + it);
  }
}

__global__ void shard_heads_kernel(const ShardPlanParams p) {
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < p.S; q += gridDim.x * blockDim.x) {
    const uint32_t k = p.okeys_sorted[q];
    const bool head = q == 0 || p.okeys_sorted[q - 1] != k;
    p.flags[q] = head ? 1u : 0u;
    if (head) atomicAdd(p.owner_counts + k / p.items_per_rank, 1u);   // integer: order-independent
  }
}

// unique recipe k (rank of its run) of owner o gets slot o*cap + (k - start[o]); its request
// (the row index AT the owner) goes to req[slot]; every row learns its slot.
__global__ void shard_fill_kernel(const ShardPlanParams p) {
  uint32_t start[9];
  start[0] = 0;
  for (int o = 0; o < p.W && o < 8; ++o) start[o + 1] = start[o] + p.owner_counts[o];
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < p.S; q += gridDim.x * blockDim.x) {
    const uint32_t key = p.okeys_sorted[q];
    const uint32_t k = p.excl[q] + p.flags[q] - 1u;
    const uint32_t o = key / p.items_per_rank;
    const uint32_t j = k - start[o];
    uint32_t slot = o * (uint32_t)p.cap + j;
    if (j >= (uint32_t)p.cap) { p.out[FR_OUT_OVERFLOW] = 2.f; slot = o * (uint32_t)p.cap; }   // reported, never OOB
    else if (p.flags[q]) p.req[slot] = (int32_t)(key - o * p.items_per_rank);
    p.slot_of_row[p.perm[q]] = (int32_t)slot;
    p.slot_sorted[q] = slot;
  }
}

static int lin_grid(int64_t n, int sm_count) {
  int64_t g = (n + 255) / 256;
  if (g > (int64_t)sm_count * 8) g = (int64_t)sm_count * 8;
  return g < 1 ? 1 : (int)g;
}
void launch_shard_prep(const ShardPlanParams& p, const Launch& l) {
  shard_prep_kernel<<<lin_grid(p.S, l.sm_count), 256, 0, l.st>>>(p); ++g_launches;
}
void launch_shard_heads(const ShardPlanParams& p, const Launch& l) {
  shard_heads_kernel<<<lin_grid(p.S, l.sm_count), 256, 0, l.st>>>(p); ++g_launches;
}
void launch_shard_fill(const ShardPlanParams& p, const Launch& l) {
  shard_fill_kernel<<<lin_grid(p.S, l.sm_count), 256, 0, l.st>>>(p); ++g_launches;
}

// owner side: requests received from every rank -> sort keys (empty slots sort last)
__global__ void serve_keys_kernel(const int32_t* __restrict__ rreq, uint32_t n, uint32_t items_per_rank,
                                  uint32_t* __restrict__ keys, uint32_t* __restrict__ n_valid) {
  uint32_t cnt = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int32_t id = rreq[i];
    keys[i] = id < 0 ? items_per_rank : (uint32_t)id;
    cnt += id >= 0;
  }
  cnt = __reduce_add_sync(FR_FULL, cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_valid, cnt);
}
void launch_serve_keys(const int32_t* rreq, uint32_t n, uint32_t items_per_rank, uint32_t* keys, uint32_t* n_valid,
                       const Launch& l) {
  cudaMemsetAsync(n_valid, 0, sizeof(uint32_t), l.st);
  serve_keys_kernel<<<lin_grid(n, l.sm_count), 256, 0, l.st>>>(rreq, n, items_per_rank, keys, n_valid); ++g_launches;
}

// one warp per requested row; an empty request yields a zero row.  With peers the row of request slot
// (source s, j) is stored straight into rank s's receive buffer over NVLink (fused gather + exchange).
__global__ void __launch_bounds__(FR_THREADS)
gather_rows_kernel(const float4* __restrict__ R, const int32_t* __restrict__ rreq, uint32_t n, int DV,
                   float4* __restrict__ out, const PeerPtrs peers, uint32_t n_table) {
  const int lane = threadIdx.x & 31;
  const uint32_t gw = blockIdx.x * FR_WARPS_PER_BLOCK + (threadIdx.x >> 5), nw = gridDim.x * FR_WARPS_PER_BLOCK;
  for (uint32_t r = gw; r < n; r += nw) {
    const int32_t id = rreq[r];
    float4* dst = out + (size_t)r * DV;
    if (peers.world > 0) {
      const uint32_t s = r / (uint32_t)peers.cap, j = r % (uint32_t)peers.cap;
      dst = peers.dst[s] + ((size_t)peers.rank * peers.cap + j) * DV;
    }
    for (int i = lane; i < DV; i += 32)
      dst[i] = (uint32_t)id < n_table ? R[(size_t)id * DV + i] : f4zero();      // (-1 = empty request; never out of the table)
  }
}
void launch_gather_rows(const float4* R, const int32_t* rreq, uint32_t n, int DV, float4* out, const PeerPtrs& peers,
                        const Launch& l, uint32_t n_table) {
  if (n == 0) return;
  int grid = (int)((n + FR_WARPS_PER_BLOCK - 1) / FR_WARPS_PER_BLOCK);
  if (grid > l.sm_count * 16) grid = l.sm_count * 16;
  gather_rows_kernel<<<grid, FR_THREADS, 0, l.st>>>(R, rreq, n, DV, out, peers, n_table); ++g_launches;
}

__global__ void add_inplace_kernel(float4* __restrict__ dst, const float4* __restrict__ src, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = add4(dst[i], src[i]);
}
void launch_add_inplace(float4* dst, const float4* src, int64_t n4, const Launch& l) {
  add_inplace_kernel<<<lin_grid(n4, l.sm_count), 256, 0, l.st>>>(dst, src, n4); ++g_launches;
}

}  // namespace fr

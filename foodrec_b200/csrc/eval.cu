// Inference + sampled evaluation kernels.
//   fwd_score_kernel        Model_Recommender.py:56-97 (model.logits)
//   eval_sampled_kernel     evaluate.py:35-66: 51 candidates/user -> dict dedup -> nlargest(K)
#include "common.cuh"
#include "internal.h"

namespace fr {

// score of one (user row in registers, item) pair; all lanes return the same value.
template <int NV>
__device__ __forceinline__ float score_pair(const float4 (&pr)[5][NV], const float4* __restrict__ Rrow,
                                            const float4 m, const float4* sCat, int DV, int lane,
                                            float a, float oma) {
  float4 rr[NV], pcs[NV];
  load_row_ro<NV>(rr, Rrow, DV, lane);
  pooled_cat<NV>(pcs, sCat, m, DV, lane);
  float hs = 0.f, ls = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float4 zs;
    zs.x = m.x * pr[1][k].x + m.y * pr[2][k].x + m.z * pr[3][k].x + m.w * pr[4][k].x;
    zs.y = m.x * pr[1][k].y + m.y * pr[2][k].y + m.z * pr[3][k].y + m.w * pr[4][k].y;
    zs.z = m.x * pr[1][k].z + m.y * pr[2][k].z + m.z * pr[3][k].z + m.w * pr[4][k].z;
    zs.w = m.x * pr[1][k].w + m.y * pr[2][k].w + m.z * pr[3][k].w + m.w * pr[4][k].w;
    hs += dot4(pr[0][k], pcs[k]);
    ls += dot4(zs, rr[k]);
  }
  hs = warp_sum(hs); ls = warp_sum(ls);
  const float n = ((m.x + m.y) + m.z) + m.w;
  return a * (hs / n) + oma * (ls / n);
}

template <int NV>
__global__ void __launch_bounds__(FR_THREADS)
fwd_score_kernel(const float4* __restrict__ P, const float4* __restrict__ R, const float4* __restrict__ Cat,
                 int DV, const int32_t* __restrict__ users, const int32_t* __restrict__ items,
                 const float4* __restrict__ cats, int cats_by_item, int n, float a, float oma,
                 float* __restrict__ scores, const HealthBlend hb) {
  extern __shared__ float4 sCat[];
  for (int i = threadIdx.x; i < 4 * DV; i += blockDim.x) sCat[i] = Cat[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * FR_WARPS_PER_BLOCK + (threadIdx.x >> 5), nw = gridDim.x * FR_WARPS_PER_BLOCK;
  int cur_u = -1;
  float4 pr[5][NV];
  // consecutive rows handled by one warp: rows [gw*RPW, (gw+1)*RPW) so that runs of the
  // same user (the 51-candidate eval feed) reuse the P row held in registers
  constexpr int RPW = 8;
  for (int blk = gw; blk * RPW < n; blk += nw) {
    for (int q = 0; q < RPW; ++q) {
      const int r = blk * RPW + q;
      if (r >= n) break;
      const int u = users[r];
      if (u != cur_u) {
#pragma unroll
        for (int s = 0; s < 5; ++s) load_row_ro<NV>(pr[s], P + ((size_t)u * 5 + s) * DV, DV, lane);
        health_blend_rows<NV>(pr, hb, u, DV, lane);
        cur_u = u;
      }
      const int it = items[r];
      const float4 m = __ldg(cats + (cats_by_item ? it : r));
      const float s = score_pair<NV>(pr, R + (size_t)it * DV, m, sCat, DV, lane, a, oma);
      if (lane == 0) scores[r] = s;
    }
  }
}

void launch_fwd_score(const ModelConsts& mc, const float4* P, const float4* R, const float4* Cat,
                      const int32_t* users, const int32_t* items, const float4* cats,
                      int cats_by_item, int n, float* scores, const HealthBlend& hb, const Launch& l) {
  if (n <= 0) return;
  int grid = (n + 8 * FR_WARPS_PER_BLOCK - 1) / (8 * FR_WARPS_PER_BLOCK);
  if (grid > l.sm_count * 8) grid = l.sm_count * 8;
  const size_t smem = (size_t)4 * mc.DV * sizeof(float4);
  ++g_launches;
  if (mc.DV <= 32)
    fwd_score_kernel<1><<<grid, FR_THREADS, smem, l.st>>>(P, R, Cat, mc.DV, users, items, cats, cats_by_item, n, mc.a, mc.oma, scores, hb);
  else
    fwd_score_kernel<2><<<grid, FR_THREADS, smem, l.st>>>(P, R, Cat, mc.DV, users, items, cats, cats_by_item, n, mc.a, mc.oma, scores, hb);
}

// One warp per test user.  Candidate j lives in lane j&31, slot j>>5 (<= 128 candidates).
constexpr int EVAL_SLOTS = 4;

template <int NV>
__global__ void __launch_bounds__(FR_THREADS)
eval_sampled_kernel(const float4* __restrict__ P, const float4* __restrict__ R, const float4* __restrict__ Cat,
                    int DV, float a, float oma, const int32_t* __restrict__ users,
                    const int32_t* __restrict__ cand, const int32_t* __restrict__ n_cand, int n_users,
                    int stride, const float4* __restrict__ cand_cats, const float4* __restrict__ item_cats,
                    int K, int32_t* __restrict__ topk_ids, int32_t* __restrict__ gt_rank,
                    float* __restrict__ scores_out, const HealthBlend hb) {
  extern __shared__ float4 sCat[];
  for (int i = threadIdx.x; i < 4 * DV; i += blockDim.x) sCat[i] = Cat[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * FR_WARPS_PER_BLOCK + (threadIdx.x >> 5), nw = gridDim.x * FR_WARPS_PER_BLOCK;
  for (int w = gw; w < n_users; w += nw) {
    const int u = users[w];
    int nc = n_cand[w];
    if (nc > stride) nc = stride;
    float4 pr[5][NV];
#pragma unroll
    for (int s = 0; s < 5; ++s) load_row_ro<NV>(pr[s], P + ((size_t)u * 5 + s) * DV, DV, lane);
    health_blend_rows<NV>(pr, hb, u, DV, lane);
    int id[EVAL_SLOTS]; float sc[EVAL_SLOTS]; bool alive[EVAL_SLOTS];
#pragma unroll
    for (int q = 0; q < EVAL_SLOTS; ++q) {
      const int j = q * 32 + lane;
      id[q] = j < nc ? cand[(size_t)w * stride + j] : -1;
      sc[q] = 0.f; alive[q] = j < nc;
    }
    for (int j = 0; j < nc; ++j) {
      int it = 0;
#pragma unroll
      for (int q = 0; q < EVAL_SLOTS; ++q) if ((j >> 5) == q) it = __shfl_sync(FR_FULL, id[q], j & 31);
      const float4 m = cand_cats ? __ldg(cand_cats + (size_t)w * stride + j) : __ldg(item_cats + it);
      const float s = score_pair<NV>(pr, R + (size_t)it * DV, m, sCat, DV, lane, a, oma);
#pragma unroll
      for (int q = 0; q < EVAL_SLOTS; ++q) if ((j >> 5) == q && (j & 31) == lane) sc[q] = s;
      if (scores_out && lane == 0) scores_out[(size_t)w * stride + j] = s;
    }
    // dict semantics (evaluate.py:60-61): the first position of an id survives and takes
    // the score of its last occurrence.
    for (int j2 = 0; j2 < nc; ++j2) {
      int idb = 0; float sb = 0.f;
#pragma unroll
      for (int q = 0; q < EVAL_SLOTS; ++q) if ((j2 >> 5) == q) {
        idb = __shfl_sync(FR_FULL, id[q], j2 & 31);
        sb = __shfl_sync(FR_FULL, sc[q], j2 & 31);
      }
#pragma unroll
      for (int q = 0; q < EVAL_SLOTS; ++q) {
        const int j = q * 32 + lane;
        if (j < nc && id[q] == idb) {
          if (j2 < j) alive[q] = false;
          else if (j2 > j) sc[q] = sb;
        }
      }
    }
    const int gt = __shfl_sync(FR_FULL, id[0], 0);
    int rank_gt = -1;
    for (int k = 0; k < K; ++k) {       // heapq.nlargest: score desc, ties -> insertion order
      float best = 0.f; int bpos = 0x7fffffff; int bid = -1;
#pragma unroll
      for (int q = 0; q < EVAL_SLOTS; ++q) {
        const int j = q * 32 + lane;
        if (alive[q] && (bpos == 0x7fffffff || sc[q] > best)) { best = sc[q]; bpos = j; bid = id[q]; }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const float ob = __shfl_xor_sync(FR_FULL, best, o);
        const int op = __shfl_xor_sync(FR_FULL, bpos, o);
        const int oi = __shfl_xor_sync(FR_FULL, bid, o);
        const bool take = (op != 0x7fffffff) && (bpos == 0x7fffffff || ob > best || (ob == best && op < bpos));
        if (take) { best = ob; bpos = op; bid = oi; }
      }
      if (bpos == 0x7fffffff) { if (lane == 0) topk_ids[(size_t)w * K + k] = -1; continue; }
#pragma unroll
      for (int q = 0; q < EVAL_SLOTS; ++q) if (q * 32 + lane == bpos) alive[q] = false;
      if (bid == gt && rank_gt < 0) rank_gt = k;
      if (lane == 0) topk_ids[(size_t)w * K + k] = bid;
    }
    if (lane == 0) gt_rank[w] = rank_gt;
  }
}

void launch_eval_sampled(const ModelConsts& mc, const float4* P, const float4* R, const float4* Cat,
                         const int32_t* users, const int32_t* cand, const int32_t* n_cand, int n_users,
                         int stride, const float4* cand_cats, const float4* item_cats, int K,
                         int32_t* topk_ids, int32_t* gt_rank, float* scores, const HealthBlend& hb, const Launch& l) {
  if (n_users <= 0) return;
  int grid = (n_users + FR_WARPS_PER_BLOCK - 1) / FR_WARPS_PER_BLOCK;
  if (grid > l.sm_count * 8) grid = l.sm_count * 8;
  const size_t smem = (size_t)4 * mc.DV * sizeof(float4);
  ++g_launches;
  if (mc.DV <= 32)
    eval_sampled_kernel<1><<<grid, FR_THREADS, smem, l.st>>>(P, R, Cat, mc.DV, mc.a, mc.oma, users, cand, n_cand, n_users,
                                                             stride, cand_cats, item_cats, K, topk_ids, gt_rank, scores, hb);
  else
    eval_sampled_kernel<2><<<grid, FR_THREADS, smem, l.st>>>(P, R, Cat, mc.DV, mc.a, mc.oma, users, cand, n_cand, n_users,
                                                             stride, cand_cats, item_cats, K, topk_ids, gt_rank, scores, hb);
}

}  // namespace fr

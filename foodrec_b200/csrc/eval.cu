// Inference + sampled evaluation kernels.
//   fwd_score_kernel        Model_Recommender.py:56-97 (model.logits)
//   eval_sampled_kernel     evaluate.py:35-66: 51 candidates/user -> dict dedup -> nlargest(K)
#include "common.cuh"
#include "internal.h"

namespace fr {

// score of one (user row in registers, item) pair; all lanes return the same value.
template <int NV, class V>
__device__ __forceinline__ float score_pair(const float4 (&pr)[5][NV], const V* __restrict__ Rrow,
                                            const float4 m, const float4* sCat, int DV, int lane,
                                            float a, float oma) {
  float4 rr[NV], pcs[NV];
  load_row_ro_t<NV>(rr, Rrow, DV, lane);
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

template <int NV, bool BF>         // BF: Personal_Memory and Recipe_Embedding are stored in bf16 (fr_set_table_format)
__global__ void __launch_bounds__(FR_THREADS)
fwd_score_kernel(const float4* __restrict__ P, const float4* __restrict__ R, const float4* __restrict__ Cat,
                 int DV, const int32_t* __restrict__ users, const int32_t* __restrict__ items,
                 const float4* __restrict__ cats, int cats_by_item, int n, float a, float oma,
                 float* __restrict__ scores, const HealthBlend hb, uint32_t n_users, uint32_t n_items) {
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
      const int it = items[r];
      // an id outside its table (tf.gather raises for it on the CPU): no row is read, the score is NaN
      if ((uint32_t)u >= n_users || (uint32_t)it >= n_items) {
        if (lane == 0) scores[r] = __int_as_float(0x7fc00000);
        continue;
      }
      if (u != cur_u) {
#pragma unroll
        for (int s = 0; s < 5; ++s) load_row_ro_t<NV>(pr[s], tab_at<BF>(P, ((size_t)u * 5 + s) * DV), DV, lane);
        health_blend_rows<NV>(pr, hb, u, DV, lane);
        cur_u = u;
      }
      const float4 m = __ldg(cats + (cats_by_item ? it : r));
      const float s = score_pair<NV>(pr, tab_at<BF>(R, (size_t)it * DV), m, sCat, DV, lane, a, oma);
      if (lane == 0) scores[r] = s;
    }
  }
}

void launch_fwd_score(const ModelConsts& mc, const float4* P, const float4* R, const float4* Cat,
                      const int32_t* users, const int32_t* items, const float4* cats,
                      int cats_by_item, int n, float* scores, const HealthBlend& hb, const Launch& l,
                      int64_t n_users, int64_t n_items, int bf16) {
  if (n <= 0) return;
  int grid = (n + 8 * FR_WARPS_PER_BLOCK - 1) / (8 * FR_WARPS_PER_BLOCK);
  if (grid > l.sm_count * 8) grid = l.sm_count * 8;
  const size_t smem = (size_t)4 * mc.DV * sizeof(float4);
  ++g_launches;
#define FR_SCORE(NVV, BFF) fwd_score_kernel<NVV, BFF><<<grid, FR_THREADS, smem, l.st>>>(P, R, Cat, mc.DV, users, items, cats, cats_by_item, \
    n, mc.a, mc.oma, scores, hb, (uint32_t)n_users, (uint32_t)n_items)
  if (mc.DV <= 32) { if (bf16) FR_SCORE(1, true); else FR_SCORE(1, false); }
  else             { if (bf16) FR_SCORE(2, true); else FR_SCORE(2, false); }
#undef FR_SCORE
}

// One warp per test user.  Candidate j's id, category weights and final score live in lane j&31, slot j>>5 (<= 128
// candidates; SLOTS = 2 covers the reference's 51); the user's five rows live in registers (lane l: 16-byte group l).
// The category term a/n * sum_c m_c <P[u,0], Cat[c]> needs the four dot products <P[u,0], Cat[c]> once per user; the
// recipe term sum_d (sum_c m_c P[u,1+c]_d) R[i]_d is computed row-per-warp with a transposed reduction (see below).
// cp.async (LDGSTS) helpers of the sampled-evaluation kernel: each lane copies ITS 16 (8: bf16) bytes of a recipe row into
// its own column of a per-warp ring in shared memory and later reads the same bytes back, so no cross-lane ordering is
// needed -- the ring is simply storage for loads in flight that does not cost registers.
__device__ __forceinline__ void cp_async_vec(float4* dst, const float4* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_vec(uint2* dst, const uint2* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ float4 ring_ld(const float4* p) { return *p; }
__device__ __forceinline__ float4 ring_ld(const uint2* p) { return bf4_to_f4(*p); }

constexpr int EV_WARPS = 4;                 // warps per CTA of eval_sampled_kernel
constexpr int EV_THREADS = EV_WARPS * 32;
constexpr int EV_CB = 8;                    // candidate rows per copy group
#ifndef FR_EV_GROUPS
#define FR_EV_GROUPS 2
#endif
#ifndef FR_EV_MINB
#define FR_EV_MINB 4
#endif
template <int NV> struct EvRing { static constexpr int GROUPS = FR_EV_GROUPS; };    // copy groups in flight per warp

// EXACT: DV == 32 * NV (D = 128, 256): every lane owns a 16-byte group of every row, no bounds predicates.
template <int NV, int SLOTS, bool BF, bool EXACT>
__global__ void __launch_bounds__(EV_THREADS, NV == 1 ? FR_EV_MINB : (FR_EV_MINB > 3 ? 3 : FR_EV_MINB))
eval_sampled_kernel(const float4* __restrict__ P, const float4* __restrict__ R, const float4* __restrict__ Cat,
                    int DV, float a, float oma, const int32_t* __restrict__ users,
                    const int32_t* __restrict__ cand, const int32_t* __restrict__ n_cand, int n_users,
                    int stride, const float4* __restrict__ cand_cats, const float4* __restrict__ item_cats,
                    int K, int32_t* __restrict__ topk_ids, int32_t* __restrict__ gt_rank,
                    float* __restrict__ scores_out, const HealthBlend hb, uint32_t n_table_users, uint32_t n_items) {
  extern __shared__ float4 smem[];
  float4* sCat = smem;                                              // [4*DV]
  for (int i = threadIdx.x; i < 4 * DV; i += blockDim.x) sCat[i] = Cat[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // per warp: the candidates' category weights and row offsets, written by the lane that owns the candidate and read
  // back by the whole warp (one broadcast LDS.128 per candidate / per four offsets instead of five shuffles)
  float4* sM = smem + 4 * DV + (threadIdx.x >> 5) * (40 * SLOTS);           // [32*SLOTS] float4
  uint32_t* sO = reinterpret_cast<uint32_t*>(sM + 32 * SLOTS);              // [32*SLOTS] uint32
  // per warp: ring of GROUPS x EV_CB row slots, [slot][k][lane] vectors -- this lane's column only is ever touched by it
  using V = typename TabVec<BF>::type;
  constexpr int GROUPS = EvRing<NV>::GROUPS;
  V* const ring = reinterpret_cast<V*>(smem + 4 * DV + EV_WARPS * (40 * SLOTS)) + ((threadIdx.x >> 5) * (GROUPS * EV_CB * NV) * 32 + lane);
  const int gw = blockIdx.x * EV_WARPS + (threadIdx.x >> 5), nw = gridDim.x * EV_WARPS;
  const typename TabVec<BF>::type* Rl = tab_at<BF>(R, 0) + lane;            // this lane's column of Recipe_Embedding
  asm volatile("" : "+l"(Rl));          // (kept as one pointer: a row address is then ONE wide multiply-add, offset * 16 + Rl)
  // The user loop is software-pipelined by one user: ids of user w+1 are loaded while user w is scored, and just before
  // user w's ranking phase (no loads in flight there) user w+1's rows are requested into L2, so that its row loads
  // find them there (half of the recipe-row reads miss L2 otherwise: the 102 MB table does not stay resident).
  constexpr int VB = (int)sizeof(typename TabVec<BF>::type);
  int un = 0, ncn = 0, idn[SLOTS];
  float4 pr[5][NV], mqn[SLOTS];                 // the NEXT user's rows / category weights while they are in flight
  auto fetch_ids = [&](int w2) {
    if (w2 < n_users) {
      un = users[w2]; ncn = n_cand[w2];
#pragma unroll
      for (int q = 0; q < SLOTS; ++q) {
        const int j = q * 32 + lane;
        idn[q] = j < stride ? cand[(size_t)w2 * stride + j] : -1;
      }
    }
  };
  // user w2's own rows (read once: streamed, evict-first, so that the recipe table, which every user re-reads, keeps
  // its place in L2) and category weights into registers, its recipe rows into L2 (hints; ids outside a table skipped)
  auto fetch_rows = [&](int w2) {
    if (w2 >= n_users) return;
    if ((uint32_t)un < n_table_users) {
#pragma unroll
      for (int s = 0; s < 5; ++s) load_row_cs_t<NV>(pr[s], tab_at<BF>(P, ((size_t)un * 5 + s) * DV), DV, lane);
    }
    const int ncl = ncn < stride ? ncn : stride;
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
      const int j = q * 32 + lane;
      mqn[q] = make_float4(1.f, 0.f, 0.f, 0.f);
      if (j < ncl && (uint32_t)idn[q] < n_items) {
        mqn[q] = cand_cats ? __ldg(cand_cats + (size_t)w2 * stride + j) : __ldg(item_cats + idn[q]);
        const char* rb = reinterpret_cast<const char*>(tab_at<BF>(R, (size_t)idn[q] * DV));
        for (int o = 0; o < DV * VB; o += 128) asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(rb + o));
      }
    }
  };
  fetch_ids(gw);
  fetch_rows(gw);
  for (int w = gw; w < n_users; w += nw) {
    const int u = un;
    int nc = ncn;
    if (nc > stride) nc = stride;
    if (nc > 32 * SLOTS) nc = 32 * SLOTS;
    int id[SLOTS]; float sc[SLOTS]; bool alive[SLOTS]; float4 mq[SLOTS];
    uint32_t roff[SLOTS];      // the candidate's row offset in 16-byte groups (a lane without a candidate -- past
                               // n_cand, or an id outside the table -- points at row 0: its score is never kept)
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
      const int j = q * 32 + lane;
      id[q] = j < nc ? idn[q] : -1;
      if ((uint32_t)id[q] >= n_items) id[q] = -1;      // a candidate outside Recipe_Embedding is dropped (never ranked)
      sc[q] = 0.f; alive[q] = id[q] >= 0;
      roff[q] = id[q] >= 0 ? (uint32_t)id[q] * (uint32_t)DV : 0u;
      mq[q] = id[q] >= 0 ? mqn[q] : make_float4(1.f, 0.f, 0.f, 0.f);
      sM[j] = mq[q]; sO[j] = roff[q];
    }
    fetch_ids(w + nw);                           // (un, ncn, idn now describe the next user)
    if ((uint32_t)u >= n_table_users) {          // a user id outside Personal_Memory: empty rank list, no row is read
      for (int k = lane; k < K; k += 32) topk_ids[(size_t)w * K + k] = -1;
      if (lane == 0) gt_rank[w] = -1;
      __syncwarp();
      fetch_rows(w + nw);
      continue;
    }
    __syncwarp();
    float b0, b1, b2, b3;
    health_blend_rows<NV>(pr, hb, u, DV, lane);
    {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        if (i < DV) {
          t0 += dot4(pr[0][k], sCat[i]); t1 += dot4(pr[0][k], sCat[DV + i]);
          t2 += dot4(pr[0][k], sCat[2 * DV + i]); t3 += dot4(pr[0][k], sCat[3 * DV + i]);
        }
      }
      b0 = warp_sum(t0); b1 = warp_sum(t1); b2 = warp_sum(t2); b3 = warp_sum(t3);
    }
    // Scoring: ROW-PER-WARP loads, one reduction per 32 candidates.  Candidate j's recipe row is read by the whole warp
    // (lane l takes 16-byte group l: a 512-byte row is 4 full lines, 4 L1 wavefronts), each lane forms its part of
    // sum_d (sum_c m_c P[u,1+c]_d) R[i]_d against the user rows it holds in registers, and the 32 partials a lane has
    // collected for a group of candidates are summed across the warp by ONE transposing butterfly (lane j ends with
    // candidate j's sum) instead of a 5-step reduction per candidate.
    // (History: the first form, row-per-warp with two shuffle reductions and two divisions per candidate, took 22 ms
    //  per 1M users; the second, lane-per-candidate -- lane j walks the row of ITS candidate, no reduction at all --
    //  7.5 ms: every load instruction touched 32 different lines for 16 bytes each, 2048 L1 wavefronts per user, and
    //  the SM's one-wavefront-per-cycle L1 port was the limit (2048 x 1M / 148 SMs / 1.9 GHz = 7.3 ms).  An eight-lanes-
    //  per-candidate variant with a reduction per four candidates measured 9.4 ms.)
    // Rows travel global -> shared memory by cp.async in groups of EV_CB = 8, GROUPS groups in flight per warp (loads in
    // flight cost no registers); group g+GROUPS is issued as soon as group g has been consumed.  Group g holds candidates
    // 8g .. 8g+7; the live groups are the first ceil(nc / 8).
    // The loop over groups is a RUNTIME loop with an 8-wide butterfly per group (7 exchanges + 2 to add the four 8-lane
    // groups), not an unrolled pass with one 32-wide butterfly per slot: fully unrolled, the kernel's hot path was ~46 KB
    // of straight-line code, more than the SM's instruction cache, and ncu showed where the time went --
    // gcc__cache_requests_type_instruction at 94 % of peak, sm__icc hit rate 86 %, the same 4.8 ms at 12, 16 or 20 warps
    // per SM, with 1-4 groups in flight, with fp32 or bf16 rows: the SMs were waiting for INSTRUCTIONS.
    constexpr int CB = EV_CB;
    const int ng = (nc + CB - 1) / CB;
    auto issue = [&](int g) {
      if (g < ng) {
        V* dst = ring + (size_t)((g % GROUPS) * CB * NV) * 32;
#pragma unroll
        for (int t = 0; t < CB; t += 4) {
          const uint4 o4 = *reinterpret_cast<const uint4*>(sO + g * CB + t);
#pragma unroll
          for (int tt = 0; tt < 4; ++tt) {
            const V* rp = Rl + (tt == 0 ? o4.x : tt == 1 ? o4.y : tt == 2 ? o4.z : o4.w);
#pragma unroll
            for (int k = 0; k < NV; ++k)
              if (EXACT || lane + 32 * k < DV) cp_async_vec(dst + ((t + tt) * NV + k) * 32, rp + 32 * k);
          }
        }
      }
      cp_async_commit();                          // (an empty group when there is nothing left: the count stays in step)
    };
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) issue(g);
    float accq[SLOTS];
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) accq[q] = 0.f;
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
      cp_async_wait<GROUPS - 1>();                // group g has landed (at most the GROUPS-1 younger ones are pending)
      const V* src = ring + (size_t)((g % GROUPS) * CB * NV) * 32;
      float v[CB];
#pragma unroll
      for (int t = 0; t < CB; ++t) {
        const float4 m = sM[g * CB + t];          // the candidate's category weights (broadcast read)
        float part = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const float4 r = (EXACT || lane + 32 * k < DV) ? ring_ld(src + (t * NV + k) * 32) : f4zero();
          float4 z;
          z.x = m.x * pr[1][k].x + m.y * pr[2][k].x + m.z * pr[3][k].x + m.w * pr[4][k].x;
          z.y = m.x * pr[1][k].y + m.y * pr[2][k].y + m.z * pr[3][k].y + m.w * pr[4][k].y;
          z.z = m.x * pr[1][k].z + m.y * pr[2][k].z + m.z * pr[3][k].z + m.w * pr[4][k].z;
          z.w = m.x * pr[1][k].w + m.y * pr[2][k].w + m.z * pr[3][k].w + m.w * pr[4][k].w;
          part += dot4(z, r);
        }
        v[t] = part;
      }
      issue(g + GROUPS);                          // reuse the slots just read
      // transposing butterfly over the low three lane bits: after the step with offset o, v[i] (i < o) is the sum over
      // 8/o lanes for the candidate whose index has this lane's bits (4, 2, 1 >= o) and low bits i; then the four
      // 8-lane groups are added: every lane ends with the complete sum of candidate 8g + (lane & 7)
#pragma unroll
      for (int off = CB / 2; off >= 1; off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
          const float send = hi ? v[i] : v[i + off];
          const float keep = hi ? v[i + off] : v[i];
          v[i] = keep + __shfl_xor_sync(FR_FULL, send, off);
        }
      }
      float tot = v[0];
      tot += __shfl_xor_sync(FR_FULL, tot, 8);
      tot += __shfl_xor_sync(FR_FULL, tot, 16);
      // candidate c lives in lane c & 31 of slot c >> 5: group g is lanes 8(g & 3) .. +7 of slot g >> 2
#pragma unroll
      for (int q = 0; q < SLOTS; ++q)
        if ((g >> 2) == q && (lane >> 3) == (g & 3)) accq[q] = tot;
    }
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
      const float4 m = mq[q];
      const float rn = __frcp_rn(((m.x + m.y) + m.z) + m.w);       // x * (1/n): exact for n = 1, 2, 4
      const float high = (((m.x * b0 + m.y * b1) + m.z * b2) + m.w * b3) * rn;   // :67-79
      const float sq = a * high + oma * (accq[q] * rn);                           // :82-96
      if (id[q] >= 0) { sc[q] = sq; if (scores_out) scores_out[(size_t)w * stride + q * 32 + lane] = sq; }
    }
    fetch_rows(w + nw);                           // the rows are dead: the next user's take their place while this one is ranked
    // dict semantics (evaluate.py:60-61): the first position of an id survives and takes the score of its last
    // occurrence.  Repeated ids are rare (0.65 % of users at 51 of 200k), so they are DETECTED first -- match.any
    // inside a slot, slot 1's candidates broadcast against slot 0 -- and the O(n) fix-up below only runs for those users.
    bool dup = true;
    if constexpr (SLOTS == 2) {
      const uint32_t m0 = __match_any_sync(FR_FULL, id[0]), m1 = __match_any_sync(FR_FULL, id[1]);
      dup = (id[0] >= 0 && m0 != (1u << lane)) || (id[1] >= 0 && m1 != (1u << lane));
      // every candidate of slot 1 (nc - 32 of them: 19 for the reference's 51) against this lane's slot-0 candidate
#pragma unroll 4
      for (int t = 0; t < nc - 32; ++t) {
        const int o = __shfl_sync(FR_FULL, id[1], t);
        dup |= (o >= 0) & (o == id[0]);
      }
      dup = __any_sync(FR_FULL, dup);
    }
    if (dup) {
      for (int j2 = 0; j2 < nc; ++j2) {
        int idb = 0; float sb = 0.f;
#pragma unroll
        for (int q = 0; q < SLOTS; ++q) if ((j2 >> 5) == q) {
          idb = __shfl_sync(FR_FULL, id[q], j2 & 31);
          sb = __shfl_sync(FR_FULL, sc[q], j2 & 31);
        }
#pragma unroll
        for (int q = 0; q < SLOTS; ++q) {
          const int j = q * 32 + lane;
          if (j < nc && id[q] == idb) {
            if (j2 < j) alive[q] = false;
            else if (j2 > j) sc[q] = sb;
          }
        }
      }
    }
    // heapq.nlargest (evaluate.py:63): score desc, ties -> insertion order.  Scores become order-preserving
    // unsigned keys (-0 folded onto +0 so that float equality is key equality); per pick: the lane's own best, one
    // REDUX max over the keys, one vote per slot for the lanes that hold it; the winning lane writes the id itself.
    uint32_t key[SLOTS];
#pragma unroll
    for (int q = 0; q < SLOTS; ++q) {
      const uint32_t b = __float_as_uint(sc[q] + 0.0f);
      key[q] = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    }
    const int gt = __shfl_sync(FR_FULL, id[0], 0);
    int rank_gt = -1;
    for (int k = 0; k < K; ++k) {
      uint32_t bk = 0u; bool any = false;
#pragma unroll
      for (int q = 0; q < SLOTS; ++q)
        if (alive[q] && (!any || key[q] > bk)) { bk = key[q]; any = true; }
      const uint32_t top = __reduce_max_sync(FR_FULL, any ? bk : 0u);
      // the winner is the LOWEST position that holds `top`: slots in order, lanes in order (two votes, no second reduce)
      int wq = -1; uint32_t wb = 0u;
#pragma unroll
      for (int q = 0; q < SLOTS; ++q) {
        const uint32_t bal = __ballot_sync(FR_FULL, alive[q] && key[q] == top);
        if (wq < 0 && bal) { wq = q; wb = bal; }
      }
      if (wq < 0) { if (lane == 0) topk_ids[(size_t)w * K + k] = -1; continue; }      // no candidate left
      const int wl = __ffs(wb) - 1;
#pragma unroll
      for (int q = 0; q < SLOTS; ++q)
        if (q == wq && lane == wl) {
          alive[q] = false;
          topk_ids[(size_t)w * K + k] = id[q];
          if (id[q] == gt && rank_gt < 0) rank_gt = k;
        }
    }
    rank_gt = (int)__reduce_max_sync(FR_FULL, (uint32_t)(rank_gt + 1)) - 1;      // (held by the lane that won that pick)
    if (lane == 0) gt_rank[w] = rank_gt;
    __syncwarp();                                 // sM / sO are rewritten for the next user
  }
}

void launch_eval_sampled(const ModelConsts& mc, const float4* P, const float4* R, const float4* Cat,
                         const int32_t* users, const int32_t* cand, const int32_t* n_cand, int n_users,
                         int stride, const float4* cand_cats, const float4* item_cats, int K,
                         int32_t* topk_ids, int32_t* gt_rank, float* scores, const HealthBlend& hb, const Launch& l,
                         int64_t n_table_users, int64_t n_items, int bf16) {
  if (n_users <= 0) return;
  const bool two = stride <= 64;
  const int NVr = mc.DV <= 32 ? 1 : 2, SL = two ? 2 : 4, VBy = bf16 ? 8 : 16;
  const int groups = NVr == 1 ? EvRing<1>::GROUPS : EvRing<2>::GROUPS;
  const size_t smem = ((size_t)4 * mc.DV + (size_t)EV_WARPS * 40 * SL) * sizeof(float4) +
                      (size_t)EV_WARPS * groups * EV_CB * NVr * 32 * VBy;
  ++g_launches;
  // launched with cudaLaunchKernelEx so that a persisting-L2 access-policy window (Recipe_Embedding: every user re-reads
  // 51 random rows of it while 2.5 KB of read-once user rows per user stream past) can ride on the LAUNCH: a stream
  // attribute does not take on the legacy default stream the Python layer hands over
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(EV_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = l.st;
  cudaLaunchAttribute attr[1];
  if (l.win) {
    attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[0].val.accessPolicyWindow = *l.win;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  const int DVv = mc.DV; const float av = mc.a, omav = mc.oma;
  const uint32_t ntu = (uint32_t)n_table_users, nit = (uint32_t)n_items;
  // persistent grid: exactly the CTAs that are resident (the user loop is software-pipelined, see the kernel)
#define FR_EVAL__(NVV, SLL, BFF, EX) do {                                                                              \
    auto kern = eval_sampled_kernel<NVV, SLL, BFF, EX>;                                                                \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                                \
    int per_sm = 1;                                                                                                    \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, EV_THREADS, smem);                                    \
    int grid = (n_users + EV_WARPS - 1) / EV_WARPS;                                                                    \
    if (grid > l.sm_count * (per_sm > 0 ? per_sm : 1)) grid = l.sm_count * (per_sm > 0 ? per_sm : 1);                  \
    cfg.gridDim = dim3(grid);                                                                                          \
    cudaLaunchKernelEx(&cfg, kern, P, R, Cat, DVv, av, omav, users, cand, n_cand, n_users, stride, cand_cats,          \
                       item_cats, K, topk_ids, gt_rank, scores, hb, ntu, nit); } while (0)
#define FR_EVAL_(NVV, SLL, BFF) do { if (mc.DV == 32 * NVV) FR_EVAL__(NVV, SLL, BFF, true); else FR_EVAL__(NVV, SLL, BFF, false); } while (0)
#define FR_EVAL(NVV, SLL) do { if (bf16) FR_EVAL_(NVV, SLL, true); else FR_EVAL_(NVV, SLL, false); } while (0)
  if (mc.DV <= 32) { if (two) FR_EVAL(1, 2); else FR_EVAL(1, 4); }
  else             { if (two) FR_EVAL(2, 2); else FR_EVAL(2, 4); }
#undef FR_EVAL__
#undef FR_EVAL_
#undef FR_EVAL
}

}  // namespace fr

// Sort-and-segment reduce of the sparse gradients, fused with the optimizer
// (Model_Recommender.py:236-240) and with Write_Memory (:106-220).
//
// Entries are sorted by key (stable radix sort, sort.cu).  A warp owns CHUNK = 32
// consecutive entries and walks the runs of equal keys inside them in batch order -- the
// order TF's unsorted_segment_sum uses.  A run fully inside the chunk is reduced and applied
// immediately (state rows are requested before the reduction so both are in flight).  A run
// that crosses a chunk boundary leaves a partial ("piece") in slot 0 (run contains the
// chunk's first entry) or slot 1 (otherwise); seg_combine_kernel then sums a crossing run's
// pieces in chunk order -- deterministic, no atomics -- and applies it once.
//
// Every kernel here is an HBM-bound row mover: one warp holds one table row at a time, a
// row of D floats moves as 16-byte vectors (lane l <-> float4 l, l+32, ...), the four
// category rows live in shared memory.
#include <cstdlib>

#include "optim.cuh"
#include "train.cuh"

namespace fr {

template <class Pol>
__global__ void __launch_bounds__(FR_THREADS)
seg_chunk_kernel(const SegCommon c, const Pol pol) {
  extern __shared__ float4 smem[];
  constexpr int NR = Pol::NR, NV = Pol::NV;
  const int DV = pol.DV();
  if (c.only_if_scaled && c.only_if_scaled[FR_OUT_SCALE] == 1.0f) return;   // single-pass step: the speculation held
  if (pol.cat_src()) {
    for (int i = threadIdx.x; i < 4 * DV; i += blockDim.x) smem[i] = pol.cat_src()[i];
    __syncthreads();
  }
  const uint32_t n = c.n_dev ? min(*c.n_dev, c.n_host) : c.n_host;
  const uint32_t nchunks = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t gw = blockIdx.x * FR_WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const uint32_t nw = gridDim.x * FR_WARPS_PER_BLOCK;
  for (uint32_t chunk = gw; chunk < nchunks; chunk += nw) {
    const uint32_t base = chunk << 5;
    const int cnt = (int)min(32u, n - base);
    const bool valid = lane < cnt;
    const uint32_t key = valid ? c.keys[base + lane] : 0xffffffffu;
    const uint32_t ent = valid ? c.perm[base + lane] : 0u;
    const uint32_t prevKey = base > 0 ? c.keys[base - 1] : 0u;
    const bool has_next = base + 32 < n;
    const uint32_t nextKey = has_next ? c.keys[base + 32] : 0u;
    const typename Pol::Entry e = pol.load_entry(ent, valid);
    const uint32_t up = __shfl_up_sync(FR_FULL, key, 1);
    const bool head = valid && (lane == 0 ? (base == 0 || prevKey != key) : (up != key));
    const uint32_t hm = __ballot_sync(FR_FULL, head);
    const uint32_t lastKey = __shfl_sync(FR_FULL, key, cnt - 1);
    const bool from_prev = !(hm & 1u);
    const bool to_next = has_next && nextKey == lastKey;
    if (c.uniq_counter && lane == 0) atomicAdd(c.uniq_counter, (uint32_t)__popc(hm));
    int e0 = 0;
#ifdef FR_PREFETCH_SEG
    {   // state rows of the second run of the chunk (the first is loaded right away)
      const uint32_t r1 = hm & ~1u;
      if (r1) pol.prefetch_state(__shfl_sync(FR_FULL, key, __ffs(r1) - 1), lane);
    }
#endif
    while (e0 < cnt) {
      const uint32_t rest = (e0 >= 31) ? 0u : (hm & ~((2u << e0) - 1u));
      const int e1 = rest ? (__ffs(rest) - 1) : cnt;
#ifdef FR_PREFETCH_SEG
      {   // ... and, while this run is processed, of the run after the next one
        const uint32_t r2 = rest & (rest - 1u);
        if (r2) pol.prefetch_state(__shfl_sync(FR_FULL, key, __ffs(r2) - 1), lane);
      }
#endif
      const bool starts = (e0 > 0) || !from_prev;
      const bool ends = (e1 < cnt) || !to_next;
      const uint32_t k = __shfl_sync(FR_FULL, key, e0);
      float4 acc[NR][NV];
#pragma unroll
      for (int s = 0; s < NR; ++s)
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[s][q] = f4zero();
      const bool contained = starts && ends;
      typename Pol::State st;
      if (contained) pol.load_state(st, k, lane);      // requested first: in flight during the reduce
      pol.accumulate(acc, e, e0, e1, lane, smem);      // (one instance: the kernel must fit the I-cache)
      if (contained) {
        pol.apply(st, k, acc, lane);
      } else {
        float4* dst = c.pieces + ((size_t)chunk * 2 + (e0 == 0 ? 0 : 1)) * NR * DV;
#pragma unroll
        for (int s = 0; s < NR; ++s)
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const int i = lane + 32 * q;
            if (i < DV) __stcg(dst + s * DV + i, acc[s][q]);
          }
      }
      e0 = e1;
    }
  }
}

template <class Pol>
__global__ void __launch_bounds__(FR_THREADS)
seg_combine_kernel(const SegCommon c, const Pol pol) {
  constexpr int NR = Pol::NR, NV = Pol::NV;
  const int DV = pol.DV();
  if (c.only_if_scaled && c.only_if_scaled[FR_OUT_SCALE] == 1.0f) return;
  const uint32_t n = c.n_dev ? min(*c.n_dev, c.n_host) : c.n_host;
  const uint32_t nchunks = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t gw = blockIdx.x * FR_WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const uint32_t nw = gridDim.x * FR_WARPS_PER_BLOCK;
  for (uint32_t chunk = gw; chunk < nchunks; chunk += nw) {
    const uint32_t base = chunk << 5;
    if (base + 32 >= n) continue;                       // last chunk: nothing continues
    const uint32_t lastKey = c.keys[base + 31];
    if (c.keys[base + 32] != lastKey) continue;         // its last run ends here
    // does the run that crosses into chunk+1 START in this chunk?
    const uint32_t key = c.keys[base + lane];
    const uint32_t prevKey = base > 0 ? c.keys[base - 1] : 0u;
    const uint32_t up = __shfl_up_sync(FR_FULL, key, 1);
    const bool head = lane == 0 ? (base == 0 || prevKey != key) : (up != key);
    const uint32_t hm = __ballot_sync(FR_FULL, head);
    if (hm == 0) continue;                              // run started in an earlier chunk
    const int s0 = 31 - __clz(hm);                      // start of the chunk's last run
    float4 acc[NR][NV];
    {
      const float4* src = c.pieces + ((size_t)chunk * 2 + (s0 == 0 ? 0 : 1)) * NR * DV;
#pragma unroll
      for (int s = 0; s < NR; ++s)
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          const int i = lane + 32 * q;
          acc[s][q] = i < DV ? __ldcg(src + s * DV + i) : f4zero();
        }
    }
    // end of the run: first position > base+31 whose key differs (binary search: the keys
    // are sorted), so the piece loads below are independent and can be pipelined
    uint32_t lo = base + 32, hi = n;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (c.keys[mid] == lastKey) lo = mid + 1; else hi = mid;
    }
    const uint32_t kend = (lo - 1) >> 5;          // last chunk holding an entry of the run
    if (c.long_list && kend - chunk >= FR_LONG_CHAIN) {   // a long chain: one block sums it (seg_combine_long_kernel)
      if (lane == 0) {
        const uint32_t w = atomicAdd(c.long_count, 1u);
        if (w < c.long_cap) c.long_list[w] = make_uint4(chunk, s0 == 0 ? 0u : 1u, kend, lastKey);
      }
      continue;
    }
    constexpr int PF = (NR * NV <= 2) ? 8 : (NR * NV <= 5 ? 4 : 2);
    for (uint32_t kc = chunk + 1; kc <= kend; kc += PF) {
      float4 buf[PF][NR][NV];
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const float4* src = c.pieces + ((size_t)(kc + u) * 2) * NR * DV;
#pragma unroll
        for (int s = 0; s < NR; ++s)
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const int i = lane + 32 * q;
            buf[u][s][q] = (kc + u <= kend && i < DV) ? __ldcg(src + s * DV + i) : f4zero();
          }
      }
#pragma unroll
      for (int u = 0; u < PF; ++u)      // summed in chunk order: deterministic
#pragma unroll
        for (int s = 0; s < NR; ++s)
#pragma unroll
          for (int q = 0; q < NV; ++q) acc[s][q] = add4(acc[s][q], buf[u][s][q]);
    }
    typename Pol::State st;
    pol.load_state(st, lastKey, lane);
    pol.apply(st, lastKey, acc, lane);
  }
}

// ---- per-entry contributions shared by the user / personal / label policies ------------
// pooledCat (normalised) and the category weights of one item row.
// One IEEE reciprocal of the category count n per entry; pooledCat and the weights w_c use
// x * (1/n) (exact for n = 1, 2, 4; 1 ulp from x/n for n = 3) -- eight divisions less per
// entry in the code the warp loops over.
template <int NV>
struct RowTerms { float4 pc[NV]; float4 w; };
template <int NV>
__device__ __forceinline__ RowTerms<NV> row_terms(const float4 m, const float4* sCat, int DV, int lane) {
  RowTerms<NV> t;
  float4 pcs[NV];
  pooled_cat<NV>(pcs, sCat, m, DV, lane);
  const float rn = __frcp_rn(((m.x + m.y) + m.z) + m.w);
  t.w = make_float4(m.x * rn, m.y * rn, m.z * rn, m.w * rn);
#pragma unroll
  for (int k = 0; k < NV; ++k) t.pc[k] = scale4(rn, pcs[k]);
  return t;
}

// ---- policy: Personal_Memory rows.  grad slice of item row r (App. A.3):
//   dP[u,0]   += g*a*pooledCat_r ;  dP[u,1+c] += g*(1-a)*w_rc*R[i_r]
// TAB: table storage (FwdParams::tab): 0 fp32, 1 bf16 P and R, 2 bf16 P + fp32 R (row-sharded step)
template <int NVV, int OPT, int TAB = 0>
struct UserPol {
  static constexpr int NV = NVV;
  static constexpr int NR = 5;
  static constexpr bool BFP = TAB != 0, BFR = TAB == 1;
  UserPolParams p;
  struct Entry { int item; float g; float4 m; };
  using State = RowState<5, NVV>;
  __device__ __forceinline__ int DV() const { return p.mc.DV; }
  __device__ __forceinline__ const float4* cat_src() const { return p.cat; }
  __device__ __forceinline__ Entry load_entry(uint32_t row, bool valid) const {
    Entry e; e.item = 0; e.g = 0.f; e.m = make_float4(1.f, 0.f, 0.f, 0.f);
    if (valid) {
      e.item = p.items[row];
      e.g = p.g[row] * p.out[FR_OUT_SCALE];
      e.m = __ldg(p.cats + (p.cats_by_item ? e.item : (int)row));
    }
    return e;
  }
  __device__ __forceinline__ void accumulate(float4 (&acc)[5][NV], const Entry& e, int e0, int e1, int lane,
                                             const float4* sCat) const {
    const int DVv = p.mc.DV;
    // two recipe rows in flight per iteration (a BPR triple is exactly one such pair)
    for (int j0 = e0; j0 < e1; j0 += 2) {
      const int jn = (j0 + 1 < e1) ? j0 + 1 : j0;
      float4 rr2[2][NV];
      load_row_ro_t<NV>(rr2[0], tab_at<BFR>(p.R, (size_t)__shfl_sync(FR_FULL, e.item, j0) * DVv), DVv, lane);
      load_row_ro_t<NV>(rr2[1], tab_at<BFR>(p.R, (size_t)__shfl_sync(FR_FULL, e.item, jn) * DVv), DVv, lane);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = j0 + u;
        if (j >= e1) break;
        const float g = __shfl_sync(FR_FULL, e.g, j);
        const float4 m = shfl4(e.m, j);
        const RowTerms<NV> t = row_terms<NV>(m, sCat, DVv, lane);
        const float ga = g * p.mc.a, go = g * p.mc.oma;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          mad4_rn(acc[0][k], ga, t.pc[k]);
          mad4_rn(acc[1][k], go * t.w.x, rr2[u][k]); mad4_rn(acc[2][k], go * t.w.y, rr2[u][k]);
          mad4_rn(acc[3][k], go * t.w.z, rr2[u][k]); mad4_rn(acc[4][k], go * t.w.w, rr2[u][k]);
        }
      }
    }
  }
  __device__ __forceinline__ void load_state(State& st, uint32_t key, int lane) const {
    fr::load_state<OPT, 5, NV, BFP>(st, p.P, p.s1, p.s2, p.last, key, p.oc, p.mc.DV, lane);
  }
  __device__ __forceinline__ void prefetch_state(uint32_t key, int lane) const {      // (FR_PREFETCH_SEG builds: fp32 tables)
    const uint32_t bytes = 5u * (uint32_t)p.mc.DV * 16u;      // a user's 5 slots are contiguous
    const size_t off = (size_t)key * 5 * p.mc.DV;
    prefetch_l2_warp(p.P + off, bytes, lane);
    if (OPT != OPT_GENERIC || p.oc.learner != FR_SGD) prefetch_l2_warp(p.s1 + off, bytes, (lane + 31) & 31);
    if (OPT != OPT_GENERIC || p.oc.learner == FR_RMSPROP) prefetch_l2_warp(p.s2 + off, bytes, (lane + 30) & 31);
  }
  __device__ __forceinline__ void apply(State& st, uint32_t key, float4 (&acc)[5][NV], int lane) const {
    apply_and_store<OPT, 5, NV, BFP>(st, p.P, p.s1, p.s2, p.last, key, acc, p.oc, p.mc.DV, lane);
  }
};

// ---- policy: the personal-memory write of Write_Memory (:149-198), personal steps only
// (first batch of epoch 0 and w.p. 1e-5: 16 mini-steps of 8 rows, Train_recommender.py:170-187).
// acc rows 0..4 = bias (sum of per-sample deltas), rows 5..9 = sum of the label-mean of
// G_old.  Runs right after the optimizer pass on P:  P = (P_opt + bias) + alpha*general_bias.
template <int NVV>
struct PersonalPol {
  static constexpr int NV = NVV;
  static constexpr int NR = 10;
  UserPolParams p;
  struct Entry { int item; float ws; float4 m; int grp; };
  struct State { float4 var[5][NVV]; };
  __device__ __forceinline__ void prefetch_state(uint32_t, int) const {}
  __device__ __forceinline__ int DV() const { return p.mc.DV; }
  __device__ __forceinline__ const float4* cat_src() const { return p.cat; }
  __device__ __forceinline__ Entry load_entry(uint32_t row, bool valid) const {
    Entry e; e.item = 0; e.ws = 0.f; e.m = make_float4(1.f, 0.f, 0.f, 0.f); e.grp = 0;
    if (valid) {
      e.item = p.items[row];
      e.ws = p.ws_row[row];
      e.m = __ldg(p.cats + (p.cats_by_item ? e.item : (int)row));
      e.grp = (int)row / p.group;
    }
    return e;
  }
  __device__ __forceinline__ void accumulate(float4 (&acc)[10][NV], const Entry& e, int e0, int e1, int lane,
                                             const float4* sCat) const {
    const int DVv = p.mc.DV;
    for (int j = e0; j < e1; ++j) {
      const int it = __shfl_sync(FR_FULL, e.item, j);
      const float ws = __shfl_sync(FR_FULL, e.ws, j);
      const float4 m = shfl4(e.m, j);
      const int grp = __shfl_sync(FR_FULL, e.grp, j);
      float4 rr[NV];
      load_row_ro<NV>(rr, p.R + (size_t)it * DVv, DVv, lane);
      const RowTerms<NV> t = row_terms<NV>(m, sCat, DVv, lane);
      const float hc = p.mc.beta_2 * ws, lc = p.mc.beta_1 * ws;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        mad4_rn(acc[0][k], hc, t.pc[k]);                                                   // :140-144
        mad4_rn(acc[1][k], lc, scale4(m.x, rr[k])); mad4_rn(acc[2][k], lc, scale4(m.y, rr[k]));   // :111-119
        mad4_rn(acc[3][k], lc, scale4(m.z, rr[k])); mad4_rn(acc[4][k], lc, scale4(m.w, rr[k]));
      }
      // (sum_l lam_l G_old[l]) / sum_l lam_l   (:170-186)
      float4 gs[5][NV];
#pragma unroll
      for (int s = 0; s < 5; ++s)
#pragma unroll
        for (int k = 0; k < NV; ++k) gs[s][k] = f4zero();
      float lsum = 0.f;
      if (p.user_labels) {
        for (int l = 0; l < p.mc.L; ++l) {
          const float lam = p.user_labels[(size_t)grp * p.mc.L + l];
          if (lam != 0.f) {
            lsum += lam;
#pragma unroll
            for (int s = 0; s < 5; ++s)
#pragma unroll
              for (int k = 0; k < NV; ++k) {
                const int i = lane + 32 * k;
                if (i < DVv) fma4(gs[s][k], lam, p.G[((size_t)l * 5 + s) * DVv + i]);
              }
          }
        }
      } else {
        const int u = p.users[grp];
        for (int q = p.lab_off[u]; q < p.lab_off[u + 1]; ++q) {
          const int l = p.lab_idx[q];
          lsum += 1.f;
#pragma unroll
          for (int s = 0; s < 5; ++s)
#pragma unroll
            for (int k = 0; k < NV; ++k) {
              const int i = lane + 32 * k;
              if (i < DVv) gs[s][k] = add4(gs[s][k], p.G[((size_t)l * 5 + s) * DVv + i]);
            }
        }
      }
#pragma unroll
      for (int s = 0; s < 5; ++s)
#pragma unroll
        for (int k = 0; k < NV; ++k) acc[5 + s][k] = add4(acc[5 + s][k], div4(gs[s][k], lsum));
    }
  }
  __device__ __forceinline__ void load_state(State& st, uint32_t key, int lane) const {
    const int DVv = p.mc.DV;
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        st.var[s][k] = i < DVv ? p.P[((size_t)key * 5 + s) * DVv + i] : f4zero();
      }
  }
  __device__ __forceinline__ void apply(State& st, uint32_t key, float4 (&acc)[10][NV], int lane) const {
    const int DVv = p.mc.DV;
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        if (i >= DVv) continue;
        float4 v = add4(st.var[s][k], acc[s][k]);                  // P + bias           (:167)
        v = add4(v, scale4(p.mc.alpha, acc[5 + s][k]));            // + alpha*general_bias (:196-198)
        p.P[((size_t)key * 5 + s) * DVv + i] = v;
      }
  }
};

// ---- policy: Recipe_Embedding rows.  dR[i] += g * z_r  (z stashed by the forward pass)
template <int NVV, int OPT, bool BFR = false>       // BFR: Recipe_Embedding stored in bf16
struct ItemPol {
  static constexpr int NV = NVV;
  static constexpr int NR = 1;
  ItemPolParams p;
  struct Entry { float g; uint32_t row; };
  using State = RowState<1, NVV>;
  __device__ __forceinline__ int DV() const { return p.mc.DV; }
  __device__ __forceinline__ const float4* cat_src() const { return nullptr; }
  __device__ __forceinline__ Entry load_entry(uint32_t row, bool valid) const {
    Entry e; e.g = 0.f; e.row = row;
    // g == nullptr: the rows are finished gradient rows received from other ranks (coef 1)
    if (valid) e.g = p.g ? p.g[row] * p.out[FR_OUT_SCALE] : 1.f;
    return e;
  }
  __device__ __forceinline__ void accumulate(float4 (&acc)[1][NV], const Entry& e, int e0, int e1, int lane,
                                             const float4*) const {
    const int DVv = p.mc.DV;
    constexpr int PF = 4;                       // z rows in flight
    for (int j0 = e0; j0 < e1; j0 += PF) {
      float4 zz[PF][NV];
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const uint32_t row = __shfl_sync(FR_FULL, e.row, min(j0 + u, e1 - 1));
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          const int i = lane + 32 * k;
          zz[u][k] = i < DVv ? __ldcg(p.z + (size_t)row * DVv + i) : f4zero();
        }
      }
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        if (j0 + u >= e1) break;
        const float g = __shfl_sync(FR_FULL, e.g, j0 + u);
#pragma unroll
        for (int k = 0; k < NV; ++k) mad4_rn(acc[0][k], g, zz[u][k]);
      }
    }
  }
  __device__ __forceinline__ void load_state(State& st, uint32_t key, int lane) const {
    fr::load_state<OPT, 1, NV, BFR>(st, p.R, p.s1, p.s2, p.last, key, p.oc, p.mc.DV, lane);
  }
  __device__ __forceinline__ void prefetch_state(uint32_t key, int lane) const {
    const uint32_t bytes = (uint32_t)p.mc.DV * 16u;
    const size_t off = (size_t)key * p.mc.DV;
    prefetch_l2_warp(p.R + off, bytes, lane);
    if (OPT != OPT_GENERIC || p.oc.learner != FR_SGD) prefetch_l2_warp(p.s1 + off, bytes, (lane + 31) & 31);
    if (OPT != OPT_GENERIC || p.oc.learner == FR_RMSPROP) prefetch_l2_warp(p.s2 + off, bytes, (lane + 30) & 31);
  }
  __device__ __forceinline__ void apply(State& st, uint32_t key, float4 (&acc)[1][NV], int lane) const {
    apply_and_store<OPT, 1, NV, BFR>(st, p.R, p.s1, p.s2, p.last, key, acc, p.oc, p.mc.DV, lane);
  }
};

// ---- policy (row-sharded training): the same reduction dR[slot] = sum g*z, but the result
// is a finished gradient row written to the send buffer of the recipe's owner instead of
// being applied (the owner applies it after the all-to-all: ItemPol with g == nullptr).
template <int NVV>
struct ItemGradPol : ItemPol<NVV, OPT_GENERIC> {
  float4* gbuf;                       // [W*cap, DV], key = slot
  PeerPtrs peers;                     // world > 0: the finished row goes straight to its owner's receive buffer
  struct State {};
  __device__ __forceinline__ void prefetch_state(uint32_t, int) const {}
  __device__ __forceinline__ void load_state(State&, uint32_t, int) const {}
  __device__ __forceinline__ void apply(State&, uint32_t key, float4 (&acc)[1][NVV], int lane) const {
    const int DVv = this->p.mc.DV;
    float4* dst = gbuf + (size_t)key * DVv;
    if (peers.world > 0) {
      const uint32_t o = key / (uint32_t)peers.cap, j = key % (uint32_t)peers.cap;
      dst = peers.dst[o] + ((size_t)peers.rank * peers.cap + j) * DVv;
    }
#pragma unroll
    for (int k = 0; k < NVV; ++k) {
      const int i = lane + 32 * k;
      if (i < DVv) dst[i] = acc[0][k];
    }
  }
};

// ---- policy: General_Memory rows (Write_Memory :201-215), entries = non-zeros of the
// label feed sorted by label:  G[l] += sum lam*ws*[beta_2*pooledCat ; beta_1*m_c*R[i]]
template <int NVV, bool BFR = false>                // BFR: the recipe rows read here are bf16
struct LabelPol {
  static constexpr int NV = NVV;
  static constexpr int NR = 5;
  LabelPolParams p;
  struct Entry { int item; float coef; float4 m; };
  struct State { float4 var[5][NVV]; };
  __device__ __forceinline__ void prefetch_state(uint32_t, int) const {}
  __device__ __forceinline__ int DV() const { return p.mc.DV; }
  __device__ __forceinline__ const float4* cat_src() const { return p.cat; }
  __device__ __forceinline__ Entry load_entry(uint32_t ent, bool valid) const {
    Entry e; e.item = 0; e.coef = 0.f; e.m = make_float4(1.f, 0.f, 0.f, 0.f);
    if (valid) {
      const uint32_t row = p.ent_row[ent];
      e.item = p.items[row];
      e.coef = p.ent_coef[ent];
      e.m = __ldg(p.cats + (p.cats_by_item ? e.item : (int)row));
    }
    return e;
  }
  // The pass is issue-bound, not bandwidth-bound (1M entries re-read 512-byte recipe rows that sit in L2), so
  // the per-entry instruction count is what matters:
  //  * slot 0 (beta_2 * coef * pooledCat) is linear in the four category rows: the per-category weights
  //    sum_j coef_j m_jc / n_j of the entries [e0, e1) are reduced across the warp ONCE per call and applied
  //    as four row FMAs, instead of a pooledCat evaluation per entry;
  //  * slots 1..4: a recipe has one or two categories, the others contribute acc + 0 -- skipped by a
  //    warp-uniform branch (bit-identical).
  __device__ __forceinline__ void accumulate(float4 (&acc)[5][NV], const Entry& e, int e0, int e1, int lane,
                                             const float4* sCat) const {
    const int DVv = p.mc.DV;
    {
      const bool in = lane >= e0 && lane < e1;
      const float s = in ? e.coef * __frcp_rn(((e.m.x + e.m.y) + e.m.z) + e.m.w) : 0.f;
      const float w0 = p.mc.beta_2 * warp_sum(s * e.m.x), w1 = p.mc.beta_2 * warp_sum(s * e.m.y);
      const float w2 = p.mc.beta_2 * warp_sum(s * e.m.z), w3 = p.mc.beta_2 * warp_sum(s * e.m.w);
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        if (i < DVv) {
          mad4_rn(acc[0][k], w0, sCat[i]); mad4_rn(acc[0][k], w1, sCat[DVv + i]);
          mad4_rn(acc[0][k], w2, sCat[2 * DVv + i]); mad4_rn(acc[0][k], w3, sCat[3 * DVv + i]);
        }
      }
    }
#ifndef FR_LABEL_PF
#define FR_LABEL_PF 8
#endif
    constexpr int PF = NV == 1 ? FR_LABEL_PF : FR_LABEL_PF / 2;   // recipe rows in flight (register budget: 128)
    for (int j0 = e0; j0 < e1; j0 += PF) {
      float4 rr4[PF][NV];
#pragma unroll
      for (int u = 0; u < PF; ++u)
        load_row_ro_t<NV>(rr4[u], tab_at<BFR>(p.R, (size_t)__shfl_sync(FR_FULL, e.item, min(j0 + u, e1 - 1)) * DVv), DVv, lane);
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int j = j0 + u;
        if (j >= e1) break;
        const float lc = p.mc.beta_1 * __shfl_sync(FR_FULL, e.coef, j);
        const float4 m = shfl4(e.m, j);
        if (m.x != 0.f) {
#pragma unroll
          for (int k = 0; k < NV; ++k) mad4_rn(acc[1][k], lc, scale4(m.x, rr4[u][k]));
        }
        if (m.y != 0.f) {
#pragma unroll
          for (int k = 0; k < NV; ++k) mad4_rn(acc[2][k], lc, scale4(m.y, rr4[u][k]));
        }
        if (m.z != 0.f) {
#pragma unroll
          for (int k = 0; k < NV; ++k) mad4_rn(acc[3][k], lc, scale4(m.z, rr4[u][k]));
        }
        if (m.w != 0.f) {
#pragma unroll
          for (int k = 0; k < NV; ++k) mad4_rn(acc[4][k], lc, scale4(m.w, rr4[u][k]));
        }
      }
    }
  }
  __device__ __forceinline__ void load_state(State& st, uint32_t key, int lane) const {
    const int DVv = p.mc.DV;
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        st.var[s][k] = i < DVv ? p.G[((size_t)key * 5 + s) * DVv + i] : f4zero();
      }
  }
  __device__ __forceinline__ void apply(State& st, uint32_t key, float4 (&acc)[5][NV], int lane) const {
    const int DVv = p.mc.DV;
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        if (i < DVv) p.G[((size_t)key * 5 + s) * DVv + i] = add4(st.var[s][k], acc[s][k]);
      }
  }
};

// ---- tiled variant for FEW, LONG runs (the label pass: <= ~100 keys, runs of thousands).
// A warp owns a tile of TC consecutive 32-entry chunks and carries its accumulator from
// chunk to chunk, so a run leaves at most two partials per TILE instead of per chunk and
// the combine chain is TC times shorter.  Same slot rule (slot 0: the run holding the
// tile's first entry, slot 1: the run holding its last), same deterministic order.
template <class Pol, int TC>
__global__ void __launch_bounds__(FR_THREADS)
seg_tile_kernel(const SegCommon c, const Pol pol) {
  extern __shared__ float4 smem[];
  constexpr int NR = Pol::NR, NV = Pol::NV;
  const int DV = pol.DV();
  if (pol.cat_src()) {
    for (int i = threadIdx.x; i < 4 * DV; i += blockDim.x) smem[i] = pol.cat_src()[i];
    __syncthreads();
  }
  const uint32_t n = c.n_dev ? min(*c.n_dev, c.n_host) : c.n_host;
  constexpr uint32_t TL = 32u * TC;
  const uint32_t ntiles = (n + TL - 1) / TL;
  const int lane = threadIdx.x & 31;
  const uint32_t gw = blockIdx.x * FR_WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const uint32_t nw = gridDim.x * FR_WARPS_PER_BLOCK;
  for (uint32_t tile = gw; tile < ntiles; tile += nw) {
    const uint32_t tbase = tile * TL;
    float4 acc[NR][NV];
#pragma unroll
    for (int s = 0; s < NR; ++s)
#pragma unroll
      for (int q = 0; q < NV; ++q) acc[s][q] = f4zero();
    bool started_in_tile = false;      // the open run's head lies inside this tile
    bool has_first = true;             // the open run holds the tile's first entry
    // the key / permutation / entry of the NEXT 32-entry chunk are requested while the current one is reduced
    // (key -> permutation -> row -> recipe id -> mask is a chain of dependent loads)
    uint32_t key_n = 0xffffffffu; typename Pol::Entry e_n;
    {
      const bool v0 = tbase + lane < n;
      key_n = v0 ? c.keys[tbase + lane] : 0xffffffffu;
      e_n = pol.load_entry(v0 ? c.perm[tbase + lane] : 0u, v0);
    }
    for (int sub = 0; sub < TC; ++sub) {
      const uint32_t base = tbase + (uint32_t)sub * 32u;
      if (base >= n) break;
      const int cnt = (int)min(32u, n - base);
      const bool valid = lane < cnt;
      const uint32_t key = key_n;
      const typename Pol::Entry e = e_n;
      if (sub + 1 < TC) {
        const uint32_t nb = base + 32u;
        const bool v1 = nb + lane < n;
        key_n = v1 ? c.keys[nb + lane] : 0xffffffffu;
        e_n = pol.load_entry(v1 ? c.perm[nb + lane] : 0u, v1);
      }
      const uint32_t prevKey = base > 0 ? c.keys[base - 1] : 0u;
      const bool has_next = base + 32 < n;
      const uint32_t nextKey = has_next ? c.keys[base + 32] : 0u;
      const uint32_t up = __shfl_up_sync(FR_FULL, key, 1);
      const bool head = valid && (lane == 0 ? (base == 0 || prevKey != key) : (up != key));
      const uint32_t hm = __ballot_sync(FR_FULL, head);
      const uint32_t lastKey = __shfl_sync(FR_FULL, key, cnt - 1);
      const bool to_next = has_next && nextKey == lastKey;
      const bool tile_ends = (sub == TC - 1) || (base + 32 >= n);
      int e0 = 0;
      while (e0 < cnt) {
        const uint32_t rest = (e0 >= 31) ? 0u : (hm & ~((2u << e0) - 1u));
        const int e1 = rest ? (__ffs(rest) - 1) : cnt;
        if ((hm >> e0) & 1u) started_in_tile = true;          // a new run begins at e0
        const uint32_t k = __shfl_sync(FR_FULL, key, e0);
        pol.accumulate(acc, e, e0, e1, lane, smem);
        const bool run_ends = (e1 < cnt) || !to_next;
        if (run_ends || tile_ends) {
          if (run_ends && started_in_tile) {                   // whole run inside the tile
            typename Pol::State st;
            pol.load_state(st, k, lane);
            pol.apply(st, k, acc, lane);
          } else {                                             // crosses a tile boundary: leave a partial
            float4* dst = c.pieces + ((size_t)tile * 2 + (has_first ? 0 : 1)) * NR * DV;
#pragma unroll
            for (int s = 0; s < NR; ++s)
#pragma unroll
              for (int q = 0; q < NV; ++q) {
                const int i = lane + 32 * q;
                if (i < DV) __stcg(dst + s * DV + i, acc[s][q]);
              }
          }
#pragma unroll
          for (int s = 0; s < NR; ++s)
#pragma unroll
            for (int q = 0; q < NV; ++q) acc[s][q] = f4zero();
          started_in_tile = false;
          has_first = false;
        }
        e0 = e1;
      }
    }
  }
}

template <class Pol, int TC>
__global__ void __launch_bounds__(FR_THREADS)
seg_tile_combine_kernel(const SegCommon c, const Pol pol) {
  constexpr int NR = Pol::NR, NV = Pol::NV;
  const int DV = pol.DV();
  const uint32_t n = c.n_dev ? min(*c.n_dev, c.n_host) : c.n_host;
  constexpr uint32_t TL = 32u * TC;
  const uint32_t ntiles = (n + TL - 1) / TL;
  const int lane = threadIdx.x & 31;
  const uint32_t gw = blockIdx.x * FR_WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const uint32_t nw = gridDim.x * FR_WARPS_PER_BLOCK;
  for (uint32_t tile = gw; tile < ntiles; tile += nw) {
    const uint32_t tbase = tile * TL, tnext = tbase + TL;
    if (tnext >= n) continue;                               // last tile: nothing continues
    const uint32_t lastKey = c.keys[tnext - 1];
    if (c.keys[tnext] != lastKey) continue;                 // its last run ends here
    // start of that run (lower bound of lastKey): this warp combines it iff the head is in this tile
    uint32_t lo = 0, hi = tnext - 1;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (c.keys[mid] < lastKey) lo = mid + 1; else hi = mid;
    }
    const uint32_t start = lo;
    if (start < tbase) continue;                            // the run began in an earlier tile
    // end of the run (upper bound)
    lo = tnext; hi = n;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (c.keys[mid] == lastKey) lo = mid + 1; else hi = mid;
    }
    const uint32_t tend = (lo - 1) / TL;                    // last tile holding an entry of the run
    if (c.long_list && tend - tile >= FR_LONG_CHAIN) {
      if (lane == 0) {
        const uint32_t w = atomicAdd(c.long_count, 1u);
        if (w < c.long_cap) c.long_list[w] = make_uint4(tile, start == tbase ? 0u : 1u, tend, lastKey);
      }
      continue;
    }
    float4 acc[NR][NV];
    {
      const float4* src = c.pieces + ((size_t)tile * 2 + (start == tbase ? 0 : 1)) * NR * DV;
#pragma unroll
      for (int s = 0; s < NR; ++s)
#pragma unroll
        for (int q = 0; q < NV; ++q) {
          const int i = lane + 32 * q;
          acc[s][q] = i < DV ? __ldcg(src + s * DV + i) : f4zero();
        }
    }
    constexpr int PF = 2;
    for (uint32_t kt = tile + 1; kt <= tend; kt += PF) {
      float4 buf[PF][NR][NV];
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const float4* src = c.pieces + ((size_t)(kt + u) * 2) * NR * DV;
#pragma unroll
        for (int s = 0; s < NR; ++s)
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const int i = lane + 32 * q;
            buf[u][s][q] = (kt + u <= tend && i < DV) ? __ldcg(src + s * DV + i) : f4zero();
          }
      }
#pragma unroll
      for (int u = 0; u < PF; ++u)      // summed in tile order: deterministic
#pragma unroll
        for (int s = 0; s < NR; ++s)
#pragma unroll
          for (int q = 0; q < NV; ++q) acc[s][q] = add4(acc[s][q], buf[u][s][q]);
    }
    typename Pol::State st;
    pol.load_state(st, lastKey, lane);
    pol.apply(st, lastKey, acc, lane);
  }
}

// ---- long chains: ONE BLOCK per crossing run.  Warp w sums a contiguous share of the run's pieces in piece
// order, the shares are then summed in warp order and the run is applied once -- a fixed summation tree that
// depends only on the chain's length, so the result is deterministic; the per-warp chain is WARPS times shorter
// (the hottest Zipf recipe of a 262,144-triple batch leaves ~940 pieces, every label of the label pass ~170).
template <class Pol>
__global__ void __launch_bounds__(FR_THREADS)
seg_combine_long_kernel(const SegCommon c, const Pol pol) {
  extern __shared__ float4 smem[];                          // [WARPS][NR*DV]
  constexpr int NR = Pol::NR, NV = Pol::NV;
  const int DV = pol.DV();
  if (c.only_if_scaled && c.only_if_scaled[FR_OUT_SCALE] == 1.0f) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t cnt = min(*c.long_count, c.long_cap);
  for (uint32_t w = blockIdx.x; w < cnt; w += gridDim.x) {
    const uint4 d = c.long_list[w];
    const uint32_t first = d.x, slot0 = d.y, last = d.z, key = d.w;
    const uint32_t per = (last - first + FR_WARPS_PER_BLOCK) / FR_WARPS_PER_BLOCK;
    const uint32_t b = first + (uint32_t)warp * per, e = min(b + per, last + 1);
    float4 acc[NR][NV];
#pragma unroll
    for (int s = 0; s < NR; ++s)
#pragma unroll
      for (int q = 0; q < NV; ++q) acc[s][q] = f4zero();
    constexpr int PF = (NR * NV <= 2) ? 8 : (NR * NV <= 5 ? 4 : 2);
    for (uint32_t k = b; k < e; k += PF) {
      float4 buf[PF][NR][NV];
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const uint32_t pi = k + u;
        const float4* src = c.pieces + ((size_t)pi * 2 + (pi == first ? slot0 : 0u)) * NR * DV;
#pragma unroll
        for (int s = 0; s < NR; ++s)
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const int i = lane + 32 * q;
            buf[u][s][q] = (pi < e && i < DV) ? __ldcg(src + s * DV + i) : f4zero();
          }
      }
#pragma unroll
      for (int u = 0; u < PF; ++u)
#pragma unroll
        for (int s = 0; s < NR; ++s)
#pragma unroll
          for (int q = 0; q < NV; ++q) acc[s][q] = add4(acc[s][q], buf[u][s][q]);
    }
#pragma unroll
    for (int s = 0; s < NR; ++s)
#pragma unroll
      for (int q = 0; q < NV; ++q) {
        const int i = lane + 32 * q;
        if (i < DV) smem[((size_t)warp * NR + s) * DV + i] = acc[s][q];
      }
    __syncthreads();
    if (warp == 0) {
#pragma unroll 1
      for (int ww = 1; ww < FR_WARPS_PER_BLOCK; ++ww)
#pragma unroll
        for (int s = 0; s < NR; ++s)
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const int i = lane + 32 * q;
            if (i < DV) acc[s][q] = add4(acc[s][q], smem[((size_t)ww * NR + s) * DV + i]);
          }
      typename Pol::State st;
      pol.load_state(st, key, lane);
      pol.apply(st, key, acc, lane);
    }
    __syncthreads();
  }
}

template <class Pol>
static void launch_combine_long(const SegCommon& c, const Pol& pol, int DV, const Launch& l) {
  if (!c.long_list) return;
  const size_t smem = (size_t)FR_WARPS_PER_BLOCK * Pol::NR * DV * sizeof(float4);
  seg_combine_long_kernel<Pol><<<l.sm_count, FR_THREADS, smem, l.st>>>(c, pol);
  ++g_launches;
}

template <class Pol, int TC>
static void launch_seg_tiled(const SegCommon& c, const Pol& pol, int DV, bool needs_cat, const Launch& l) {
  const uint32_t ntiles = (c.n_host + 32 * TC - 1) / (32 * TC);
  if (ntiles == 0) return;
  int grid = (int)((ntiles + FR_WARPS_PER_BLOCK - 1) / FR_WARPS_PER_BLOCK);
  const size_t smem = needs_cat ? (size_t)4 * DV * sizeof(float4) : 0;
  const int cap = l.sm_count * 16;
  if (grid > cap) grid = cap;
  seg_tile_kernel<Pol, TC><<<grid, FR_THREADS, smem, l.st>>>(c, pol);
  if (l.mid) cudaEventRecord(l.mid, l.st);
  seg_tile_combine_kernel<Pol, TC><<<grid, FR_THREADS, 0, l.st>>>(c, pol);
  g_launches += 2;
  launch_combine_long(c, pol, DV, l);
}

template <class Pol>
static void launch_seg(const SegCommon& c, const Pol& pol, int DV, bool needs_cat, const Launch& l) {
  const uint32_t nchunks = (c.n_host + 31) / 32;
  if (nchunks == 0) return;
  int grid = (int)((nchunks + FR_WARPS_PER_BLOCK - 1) / FR_WARPS_PER_BLOCK);
  const size_t smem = needs_cat ? (size_t)4 * DV * sizeof(float4) : 0;
  // more CTAs than are resident: the block scheduler then balances the uneven per-chunk
  // work (measured: x16 beats a resident-sized persistent grid)
  const int cap = l.sm_count * 16;
  if (grid > cap) grid = cap;
  seg_chunk_kernel<Pol><<<grid, FR_THREADS, smem, l.st>>>(c, pol);
  if (l.mid) cudaEventRecord(l.mid, l.st);
  seg_combine_kernel<Pol><<<grid, FR_THREADS, 0, l.st>>>(c, pol);
  g_launches += 2;
  launch_combine_long(c, pol, DV, l);
}

// =====================================================================================================
// SINGLE-PASS training of Personal_Memory (lazy Adam): forward + loss + norms + segment reduce + optimizer in ONE
// walk over the user-sorted item rows.
//
// The two-pass step reads P, m, v of every batch user twice: the forward needs the caught-up row to score it, the
// update pass needs it again because tf.clip_by_global_norm (Model_Recommender.py:237) sits between the gradient and
// apply_gradients -- the scale is only known once EVERY sample has been scored.  ncu, round 1: 2.45 GB + 3.65 GB of
// DRAM traffic for the two kernels, 1.03 ms of a 1.66 ms step.  But the clip is inactive unless the global norm exceeds
// 5 (never, at a mean over 262,144 triples), and then the scale is exactly 1.0f.  So this kernel SPECULATES scale = 1:
// a warp takes a run of sorted rows of one user, loads the user's P, m, v once, catches them up (lazy Adam), scores
// the user's samples (same arithmetic as fwd_train_kernel: loss, g, per-slice norms, dCat partials, the z stash for
// the recipe pass), reduces the gradient slices in batch order (same arithmetic as UserPol) and applies Adam.
// Nothing is overwritten: the new rows go to the OTHER copy of a double-buffered table (fr_set_shadow), and a row's
// stamp carries one bit saying which copy is current.  Once finalize_kernel knows the norm, user_commit_kernel flips
// the bits of the batch's users if the scale is exactly 1 (a 2 MB pass); otherwise the bits stay, the speculative rows
// are simply never looked at, and the ordinary update pass runs with the true scale (api.cu).  Runs that cross a
// 32-row chunk leave partial sums and are applied by seg_combine_kernel<FusedUserPol>, same protocol.
template <int NVV, int OPT>
struct FusedUserPol {               // load_state / apply for the crossing runs (seg_combine_kernel, seg_combine_long_kernel)
  static constexpr int NV = NVV;
  static constexpr int NR = 5;
  FusedParams p;
  struct State { float4 var[5][NVV], s1[5][NVV], s2[5][NVV]; int last, sh; };
  __device__ __forceinline__ int DV() const { return p.mc.DV; }
  __device__ __forceinline__ void load_state(State& st, uint32_t key, int lane) const {
    const int l = p.last[key];
    st.sh = (l >> 30) & 1; st.last = l & 0x3fffffff;
    const int DVv = p.mc.DV;
    const size_t base = (size_t)key * 5 * DVv;
    const float4 *Ps = p.P[st.sh] + base, *ms = p.m[st.sh] + base, *vs = p.v[st.sh] + base;
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        const bool ok = i < DVv;
        st.var[s][k] = ok ? __ldcs(Ps + s * DVv + i) : f4zero();
        st.s1[s][k] = ok ? __ldcs(ms + s * DVv + i) : f4zero();
        st.s2[s][k] = ok ? __ldcs(vs + s * DVv + i) : f4zero();
      }
  }
  // catch-up to step-1 (already done for a run the fused kernel scored itself: caught == true), Adam with the summed
  // slices, new rows into the OTHER copy; the stamp is left to user_commit_kernel
  __device__ __forceinline__ void apply_to_other(State& st, uint32_t key, const float4 (&grad)[5][NV], int lane, bool caught) const {
    if (!caught) adam_catchup<OPT, 5 * NV>(&st.var[0][0], &st.s1[0][0], &st.s2[0][0], st.last, p.oc.step - 1, p.oc);
    const int DVv = p.mc.DV;
    const size_t base = (size_t)key * 5 * DVv;
    float4 *Pd = p.P[st.sh ^ 1] + base, *md = p.m[st.sh ^ 1] + base, *vd = p.v[st.sh ^ 1] + base;
#pragma unroll
    for (int s = 0; s < 5; ++s)
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        if (i >= DVv) continue;
        float4 var = st.var[s][k], a = st.s1[s][k], b = st.s2[s][k];
        const float4 g = grad[s][k];
        adam_touch(var.x, a.x, b.x, g.x, p.oc); adam_touch(var.y, a.y, b.y, g.y, p.oc);
        adam_touch(var.z, a.z, b.z, g.z, p.oc); adam_touch(var.w, a.w, b.w, g.w, p.oc);
        __stcs(Pd + s * DVv + i, var); __stcs(md + s * DVv + i, a); __stcs(vd + s * DVv + i, b);
      }
  }
  __device__ __forceinline__ void apply(State& st, uint32_t key, float4 (&acc)[5][NV], int lane) const {
    apply_to_other(st, key, acc, lane, false);
  }
};

template <int NV, int GROUP, int OPT>
__global__ void __launch_bounds__(FR_THREADS, NV == 1 ? 2 : 1)      // D <= 128: two CTAs per SM (128 registers)
user_fused_kernel(const SegCommon c, const FusedUserPol<NV, OPT> pol) {
  extern __shared__ float4 smem[];
  const FusedParams& p = pol.p;
  const int DV = p.mc.DV;
  float4* sCat = smem;                          // [4*DV]
  float4* sgc = smem + 4 * DV;                  // [WARPS][4*DV]: dCat partial of each warp
  __shared__ float red_loss[FR_WARPS_PER_BLOCK], red_nrm[FR_WARPS_PER_BLOCK];
  for (int i = threadIdx.x; i < 4 * DV; i += blockDim.x) sCat[i] = p.cat[i];
  for (int i = threadIdx.x; i < FR_WARPS_PER_BLOCK * 4 * DV; i += blockDim.x) sgc[i] = f4zero();
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4* mygc = sgc + warp * 4 * DV;
  const uint32_t n = c.n_host;
  const uint32_t nchunks = (n + 31) >> 5;
  const uint32_t gw = blockIdx.x * FR_WARPS_PER_BLOCK + warp, nw = gridDim.x * FR_WARPS_PER_BLOCK;
  const float a = p.a, oma = p.oma, Bf = p.Bnorm;
  float lossacc = 0.f, nrmacc = 0.f;
  for (uint32_t chunk = gw; chunk < nchunks; chunk += nw) {
    const uint32_t base = chunk << 5;
    const int cnt = (int)min(32u, n - base);
    const bool valid = lane < cnt;
    const uint32_t key = valid ? c.keys[base + lane] : 0xffffffffu;
    const uint32_t ent = valid ? c.perm[base + lane] : 0u;
    const uint32_t prevKey = base > 0 ? c.keys[base - 1] : 0u;
    const bool has_next = base + 32 < n;
    const uint32_t nextKey = has_next ? c.keys[base + 32] : 0u;
    // per-lane entry data: recipe id, category mask, label, and the stamp of the lane's user
    int e_item = 0; float4 e_m = make_float4(1.f, 0.f, 0.f, 0.f); float e_y = 0.f; int e_last = 0;
    if (valid) {
      e_item = p.items[ent];
      e_m = __ldg(p.cats + (p.cats_by_item ? e_item : (int)ent));
      if (GROUP == 1) e_y = p.labels[ent];
      e_last = p.last[key];
    }
    const uint32_t up = __shfl_up_sync(FR_FULL, key, 1);
    const bool head = valid && (lane == 0 ? (base == 0 || prevKey != key) : (up != key));
    const uint32_t hm = __ballot_sync(FR_FULL, head);
    const uint32_t lastKey = __shfl_sync(FR_FULL, key, cnt - 1);
    const bool from_prev = !(hm & 1u);
    const bool to_next = has_next && nextKey == lastKey;
    if (c.uniq_counter && lane == 0) atomicAdd(c.uniq_counter, (uint32_t)__popc(hm));
    // (Measured and rejected: TMA bulk prefetches of the NEXT runs' P / m / v into L2 -- the kernel is latency-bound,
    //  16 warps per SM each holding one user's state in 60 registers -- made it slower, 0.88 vs 0.80 ms: the prefetch
    //  traffic competes with the demand loads for the same DRAM queues; same finding as FR_PREFETCH_SEG in round 1.)
    int e0 = 0;
    while (e0 < cnt) {
      const uint32_t rest = (e0 >= 31) ? 0u : (hm & ~((2u << e0) - 1u));
      const int e1 = rest ? (__ffs(rest) - 1) : cnt;
      const bool contained = ((e0 > 0) || !from_prev) && ((e1 < cnt) || !to_next);
      const uint32_t k = __shfl_sync(FR_FULL, key, e0);
      typename FusedUserPol<NV, OPT>::State st;
      {
        const int l = __shfl_sync(FR_FULL, e_last, e0);
        st.sh = (l >> 30) & 1; st.last = l & 0x3fffffff;
        const size_t ub = (size_t)k * 5 * DV;
        const float4 *Ps = p.P[st.sh] + ub, *ms = p.m[st.sh] + ub, *vs = p.v[st.sh] + ub;
#pragma unroll
        for (int s = 0; s < 5; ++s)
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const int i = lane + 32 * q;
            const bool ok = i < DV;
            st.var[s][q] = ok ? __ldcs(Ps + s * DV + i) : f4zero();
            st.s1[s][q] = ok ? __ldcs(ms + s * DV + i) : f4zero();
            st.s2[s][q] = ok ? __ldcs(vs + s * DV + i) : f4zero();
          }
      }
      // the recipe rows of the run's first group do not depend on the user's state: requested before the catch-up
      // arithmetic so that both are in flight together (the next group's rows are requested while this one is scored)
      float4 rr[GROUP][NV];
#pragma unroll
      for (int j = 0; j < GROUP; ++j)
        load_row_ro<NV>(rr[j], p.R + (size_t)__shfl_sync(FR_FULL, e_item, e0 + j) * DV, DV, lane);
      // the rows TF would see at this step: decay-only steps the user sat out
      adam_catchup<OPT, 5 * NV>(&st.var[0][0], &st.s1[0][0], &st.s2[0][0], st.last, p.oc.step - 1, p.oc);
      float4 acc[5][NV];
#pragma unroll
      for (int s = 0; s < 5; ++s)
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[s][q] = f4zero();
      for (int j0 = e0; j0 < e1; j0 += GROUP) {
        float4 pcn[GROUP][NV];
        float4 mm[GROUP]; float sc[GROUP], rnn[GROUP], nzq[GROUP], nRq[GROUP], npcq[GROUP];
        uint32_t row[GROUP];
#pragma unroll
        for (int j = 0; j < GROUP; ++j) {
          row[j] = __shfl_sync(FR_FULL, ent, j0 + j);
          mm[j] = shfl4(e_m, j0 + j);
        }
#pragma unroll
        for (int j = 0; j < GROUP; ++j) {           // Model_Recommender.py:56-97, as fwd_train_kernel
          const float4 m = mm[j];
          const float nn = ((m.x + m.y) + m.z) + m.w;
          const float rn = __frcp_rn(nn);
          float4 pcs[NV], zs[NV];
          pooled_cat<NV>(pcs, sCat, m, DV, lane);
          float hs = 0.f, ls = 0.f, nz = 0.f, nR = 0.f, npc = 0.f;
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            zs[q].x = m.x * st.var[1][q].x + m.y * st.var[2][q].x + m.z * st.var[3][q].x + m.w * st.var[4][q].x;
            zs[q].y = m.x * st.var[1][q].y + m.y * st.var[2][q].y + m.z * st.var[3][q].y + m.w * st.var[4][q].y;
            zs[q].z = m.x * st.var[1][q].z + m.y * st.var[2][q].z + m.z * st.var[3][q].z + m.w * st.var[4][q].z;
            zs[q].w = m.x * st.var[1][q].w + m.y * st.var[2][q].w + m.z * st.var[3][q].w + m.w * st.var[4][q].w;
            hs += dot4(st.var[0][q], pcs[q]);
            ls += dot4(zs[q], rr[j][q]);
            nz += dot4(zs[q], zs[q]);
            if constexpr (GROUP == 1) { nR += dot4(rr[j][q], rr[j][q]); npc += dot4(pcs[q], pcs[q]); }
          }
          hs = warp_sum(hs); ls = warp_sum(ls);
          sc[j] = a * (hs * rn) + oma * (ls * rn);
          const float zc = oma * rn;
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const int i = lane + 32 * q;
            if (i < DV) __stcg(p.z + (size_t)row[j] * DV + i, scale4(zc, zs[q]));
            pcn[j][q] = scale4(rn, pcs[q]);
          }
          rnn[j] = rn;
          const float inv2 = rn * rn;
          nzq[j] = nz * inv2; nRq[j] = nR * inv2; npcq[j] = npc * inv2;
        }
        float gq[GROUP];
        if constexpr (GROUP == 1) {
          const float s = sc[0], y = __shfl_sync(FR_FULL, e_y, j0);
          const float e = expf(-fabsf(s));
          const float loss = fmaxf(s, 0.f) - s * y + log1pf(e);
          const float sig = s >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
          const float g = (sig - y) / Bf;
          const float4 m = mm[0];
          const float sumsq_m = m.x * m.x + m.y * m.y + m.z * m.z + m.w * m.w;
          const float nrm = g * g * (a * a * npcq[0] + oma * oma * (nRq[0] * sumsq_m + nzq[0]));
          const float ga = g * a, rn = rnn[0];
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const int i = lane + 32 * q;
            if (i < DV) {
              float4 t;
              t = mygc[i]; fma4(t, ga * (m.x * rn), st.var[0][q]); mygc[i] = t;
              t = mygc[DV + i]; fma4(t, ga * (m.y * rn), st.var[0][q]); mygc[DV + i] = t;
              t = mygc[2 * DV + i]; fma4(t, ga * (m.z * rn), st.var[0][q]); mygc[2 * DV + i] = t;
              t = mygc[3 * DV + i]; fma4(t, ga * (m.w * rn), st.var[0][q]); mygc[3 * DV + i] = t;
            }
          }
          lossacc += loss; nrmacc += nrm;
          gq[0] = g;
          if (lane == 0) { p.g[row[0]] = g; p.scores[row[0]] = s; }
        } else {
          const float s = sc[0] - sc[GROUP - 1];
          const float e = expf(-fabsf(s));
          const float loss = fmaxf(s, 0.f) - s + log1pf(e);
          const float sig = s >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
          const float h = (sig - 1.f) / Bf;
          const float4 m0 = mm[0], m1 = mm[GROUP - 1];
          const float4 w0 = scale4(rnn[0], m0), w1 = scale4(rnn[GROUP - 1], m1);
          float dq = 0.f;
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const float4 r0 = rr[0][q], r1 = rr[GROUP - 1][q];
            float4 d = make_float4(pcn[0][q].x - pcn[GROUP - 1][q].x, pcn[0][q].y - pcn[GROUP - 1][q].y,
                                   pcn[0][q].z - pcn[GROUP - 1][q].z, pcn[0][q].w - pcn[GROUP - 1][q].w);
            dq += a * a * dot4(d, d);
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const float wa = comp(w0, cc), wb = comp(w1, cc);
              d = make_float4(wa * r0.x - wb * r1.x, wa * r0.y - wb * r1.y, wa * r0.z - wb * r1.z, wa * r0.w - wb * r1.w);
              dq += oma * oma * dot4(d, d);
            }
          }
          const float nrm = h * h * (dq + oma * oma * (nzq[0] + nzq[GROUP - 1]));
          const float ha = h * a;
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const int i = lane + 32 * q;
            if (i < DV) {
              float4 t;
              t = mygc[i]; fma4(t, ha * (w0.x - w1.x), st.var[0][q]); mygc[i] = t;
              t = mygc[DV + i]; fma4(t, ha * (w0.y - w1.y), st.var[0][q]); mygc[DV + i] = t;
              t = mygc[2 * DV + i]; fma4(t, ha * (w0.z - w1.z), st.var[0][q]); mygc[2 * DV + i] = t;
              t = mygc[3 * DV + i]; fma4(t, ha * (w0.w - w1.w), st.var[0][q]); mygc[3 * DV + i] = t;
            }
          }
          lossacc += loss; nrmacc += nrm;
          gq[0] = h; gq[GROUP - 1] = -h;
          if (lane == 0) {
            p.g[row[0]] = h; p.g[row[GROUP - 1]] = -h;
            p.scores[row[0]] = sc[0]; p.scores[row[GROUP - 1]] = sc[GROUP - 1];
          }
        }
        // gradient slices of this group's rows, in batch order, rounded like UserPol::accumulate (clip scale 1)
#pragma unroll
        for (int j = 0; j < GROUP; ++j) {
          const float4 wj = scale4(rnn[j], mm[j]);
          const float ga = gq[j] * a, go = gq[j] * oma;
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            mad4_rn(acc[0][q], ga, pcn[j][q]);
            mad4_rn(acc[1][q], go * wj.x, rr[j][q]); mad4_rn(acc[2][q], go * wj.y, rr[j][q]);
            mad4_rn(acc[3][q], go * wj.z, rr[j][q]); mad4_rn(acc[4][q], go * wj.w, rr[j][q]);
          }
        }
        if (j0 + GROUP < e1) {
#pragma unroll
          for (int j = 0; j < GROUP; ++j)
            load_row_ro<NV>(rr[j], p.R + (size_t)__shfl_sync(FR_FULL, e_item, j0 + GROUP + j) * DV, DV, lane);
        }
      }
      if (contained) {
        pol.apply_to_other(st, k, acc, lane, true);
      } else {
        float4* dst = c.pieces + ((size_t)chunk * 2 + (e0 == 0 ? 0 : 1)) * 5 * DV;
#pragma unroll
        for (int s = 0; s < 5; ++s)
#pragma unroll
          for (int q = 0; q < NV; ++q) {
            const int i = lane + 32 * q;
            if (i < DV) __stcg(dst + s * DV + i, acc[s][q]);
          }
      }
      e0 = e1;
    }
  }
  // deterministic block reduction (fixed warp order), one partial per block -- as fwd_train_kernel
  nrmacc = warp_sum(nrmacc);
  if (lane == 0) { red_loss[warp] = lossacc; red_nrm[warp] = nrmacc; }
  __syncthreads();
  for (int j = threadIdx.x; j < 4 * DV; j += blockDim.x) {
    float4 s = sgc[j];
#pragma unroll
    for (int w = 1; w < FR_WARPS_PER_BLOCK; ++w) s = add4(s, sgc[w * 4 * DV + j]);
    p.part_gcat[(size_t)blockIdx.x * 4 * DV + j] = s;
  }
  if (threadIdx.x == 0) {
    float l = 0.f, q = 0.f;
#pragma unroll
    for (int w = 0; w < FR_WARPS_PER_BLOCK; ++w) { l += red_loss[w]; q += red_nrm[w]; }
    p.part_loss[blockIdx.x] = l; p.part_nrm[blockIdx.x] = q;
  }
}

// the speculation held (scale is exactly 1): the rows written to the other copy become current
__global__ void __launch_bounds__(256)
user_commit_kernel(const uint32_t* __restrict__ keys, uint32_t n, int32_t* __restrict__ last, const float* __restrict__ out, int step) {
  if (out[FR_OUT_SCALE] != 1.0f) return;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t k = keys[i];
    if (i == 0 || keys[i - 1] != k) {
      const int l = last[k];
      last[k] = step | ((l ^ (int)FR_SHADOW_BIT) & (int)FR_SHADOW_BIT);
    }
  }
}

// Rows whose current copy is the shadow go back to the caller's tables and their bit is cleared.  pred != nullptr:
// only if the step's clip scale is NOT exactly 1 (the speculation failed: the ordinary update pass follows, which
// works on the caller's tables).
__global__ void __launch_bounds__(FR_THREADS)
shadow_consolidate_kernel(int32_t* __restrict__ last, int64_t n_users, int rowDV, float4* __restrict__ P, float4* __restrict__ m,
                          float4* __restrict__ v, const float4* __restrict__ Pa, const float4* __restrict__ ma,
                          const float4* __restrict__ va, const float* __restrict__ pred) {
  if (pred && pred[FR_OUT_SCALE] == 1.0f) return;
  const int lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * FR_WARPS_PER_BLOCK + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * FR_WARPS_PER_BLOCK;
  for (int64_t u0 = gw * 32; u0 < n_users; u0 += nw * 32) {
    const int64_t u = u0 + lane;
    const int l = u < n_users ? last[u] : 0;
    uint32_t mask = __ballot_sync(FR_FULL, (l & (int)FR_SHADOW_BIT) != 0);
    if (l & (int)FR_SHADOW_BIT) last[u] = l & ~(int)FR_SHADOW_BIT;
    while (mask) {
      const int j = __ffs(mask) - 1;
      mask &= mask - 1;
      const size_t b = (size_t)(u0 + j) * rowDV;
      for (int i = lane; i < rowDV; i += 32) {
        __stcs(P + b + i, __ldcs(Pa + b + i)); __stcs(m + b + i, __ldcs(ma + b + i)); __stcs(v + b + i, __ldcs(va + b + i));
      }
    }
  }
}

void launch_shadow_consolidate(int32_t* last, int64_t n_users, int rowDV, float4* P, float4* m, float4* v, const float4* Pa,
                               const float4* ma, const float4* va, const float* pred, const Launch& l) {
  if (n_users <= 0) return;
  int64_t grid = (n_users + FR_THREADS - 1) / FR_THREADS;
  if (grid > (int64_t)l.sm_count * 8) grid = (int64_t)l.sm_count * 8;
  shadow_consolidate_kernel<<<(int)grid, FR_THREADS, 0, l.st>>>(last, n_users, rowDV, P, m, v, Pa, ma, va, pred);
  ++g_launches;
}

int user_fused_grid(uint32_t n_rows, int sm_count) {
  const uint32_t nchunks = (n_rows + 31) / 32;
  int grid = (int)((nchunks + FR_WARPS_PER_BLOCK - 1) / FR_WARPS_PER_BLOCK);
  // exactly the CTAs that are resident (two per SM): with a static grid-stride assignment a second, partial wave
  // would leave SMs idle at the end (4 per SM measured 0.83 ms: 592 CTAs in two waves, 3 or 4 chunks per warp);
  // FOODREC_FUSED_CTAS_PER_SM overrides (<= 4: the dCat / loss / norm partial workspace holds 4 per SM)
  static int per_sm = -1;
  if (per_sm < 0) { const char* e = getenv("FOODREC_FUSED_CTAS_PER_SM"); per_sm = e ? atoi(e) : 2; if (per_sm < 1 || per_sm > 4) per_sm = 2; }
  const int cap = sm_count * per_sm;
  if (grid > cap) grid = cap;
  return grid < 1 ? 1 : grid;
}

void launch_user_fused(int NV, int group, const SegCommon& c, const FusedParams& p, int grid, const Launch& l) {
  const int opt = opt_of(p.oc.learner, p.oc.adam_mode);
  const size_t smem = (size_t)(4 * p.mc.DV) * sizeof(float4) * (1 + FR_WARPS_PER_BLOCK);
  const uint32_t nchunks = (c.n_host + 31) / 32;
  int cgrid = (int)((nchunks + FR_WARPS_PER_BLOCK - 1) / FR_WARPS_PER_BLOCK);
  if (cgrid > l.sm_count * 16) cgrid = l.sm_count * 16;
#define FR_FUSED(NVV, GG, OO) do { FusedUserPol<NVV, OO> pol{p};                                            \
    user_fused_kernel<NVV, GG, OO><<<grid, FR_THREADS, smem, l.st>>>(c, pol);                                \
    if (l.mid) cudaEventRecord(l.mid, l.st);                                                                 \
    seg_combine_kernel<FusedUserPol<NVV, OO>><<<cgrid, FR_THREADS, 0, l.st>>>(c, pol); } while (0)
#define FR_FUSED_O(NVV, GG) do { if (opt == OPT_ADAM_EXACT) FR_FUSED(NVV, GG, OPT_ADAM_EXACT); else FR_FUSED(NVV, GG, OPT_ADAM_SERIES); } while (0)
  if (NV == 1) { if (group == 1) FR_FUSED_O(1, 1); else FR_FUSED_O(1, 2); }
  else         { if (group == 1) FR_FUSED_O(2, 1); else FR_FUSED_O(2, 2); }
#undef FR_FUSED_O
#undef FR_FUSED
  g_launches += 2;
}

void launch_user_commit(const uint32_t* keys, uint32_t n, int32_t* last, const float* out, int step, const Launch& l) {
  int grid = (int)((n + 255) / 256);
  if (grid > l.sm_count * 8) grid = l.sm_count * 8;
  if (grid < 1) grid = 1;
  user_commit_kernel<<<grid, 256, 0, l.st>>>(keys, n, last, out, step);
  ++g_launches;
}

#define FR_DISPATCH_NV_OPT(NVx, OPTx, ...)                                                  \
  do {                                                                                        \
    if ((NVx) == 1) {                                                                         \
      switch (OPTx) {                                                                         \
        case OPT_ADAM_DENSE:  { constexpr int NV_ = 1, OPT_ = OPT_ADAM_DENSE;  __VA_ARGS__ } break;  \
        case OPT_ADAM_EXACT:  { constexpr int NV_ = 1, OPT_ = OPT_ADAM_EXACT;  __VA_ARGS__ } break;  \
        case OPT_ADAM_SERIES: { constexpr int NV_ = 1, OPT_ = OPT_ADAM_SERIES; __VA_ARGS__ } break;  \
        default:              { constexpr int NV_ = 1, OPT_ = OPT_GENERIC;     __VA_ARGS__ } break;  \
      }                                                                                       \
    } else {                                                                                  \
      switch (OPTx) {                                                                         \
        case OPT_ADAM_DENSE:  { constexpr int NV_ = 2, OPT_ = OPT_ADAM_DENSE;  __VA_ARGS__ } break;  \
        case OPT_ADAM_EXACT:  { constexpr int NV_ = 2, OPT_ = OPT_ADAM_EXACT;  __VA_ARGS__ } break;  \
        case OPT_ADAM_SERIES: { constexpr int NV_ = 2, OPT_ = OPT_ADAM_SERIES; __VA_ARGS__ } break;  \
        default:              { constexpr int NV_ = 2, OPT_ = OPT_GENERIC;     __VA_ARGS__ } break;  \
      }                                                                                       \
    }                                                                                         \
  } while (0)

void launch_user_pass(int NV, const SegCommon& c, const UserPolParams& p, const Launch& l) {
  const int opt = opt_of(p.oc.learner, p.oc.adam_mode);
  if (p.tab != 0) {      // bf16 tables: SGD / Adagrad / RMSProp only (fr_set_table_format checks the learner)
    if (NV == 1) { if (p.tab == 1) { UserPol<1, OPT_GENERIC, 1> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); }
                   else { UserPol<1, OPT_GENERIC, 2> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); } }
    else         { if (p.tab == 1) { UserPol<2, OPT_GENERIC, 1> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); }
                   else { UserPol<2, OPT_GENERIC, 2> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); } }
    return;
  }
  FR_DISPATCH_NV_OPT(NV, opt, { UserPol<NV_, OPT_> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); });
}
void launch_personal_pass(int NV, const SegCommon& c, const UserPolParams& p, const Launch& l) {
  if (NV == 1) { PersonalPol<1> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); }
  else { PersonalPol<2> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); }
}
void launch_item_pass(int NV, const SegCommon& c, const ItemPolParams& p, const Launch& l) {
  const int opt = opt_of(p.oc.learner, p.oc.adam_mode);
  if (p.tab != 0) {
    if (NV == 1) { ItemPol<1, OPT_GENERIC, true> pol{p}; launch_seg(c, pol, p.mc.DV, false, l); }
    else { ItemPol<2, OPT_GENERIC, true> pol{p}; launch_seg(c, pol, p.mc.DV, false, l); }
    return;
  }
  FR_DISPATCH_NV_OPT(NV, opt, { ItemPol<NV_, OPT_> pol{p}; launch_seg(c, pol, p.mc.DV, false, l); });
}
void launch_item_grad_pass(int NV, const SegCommon& c, const ItemPolParams& p, float4* gbuf, const PeerPtrs& peers,
                           const Launch& l) {
  if (NV == 1) { ItemGradPol<1> pol{{p}, gbuf, peers}; launch_seg(c, pol, p.mc.DV, false, l); }
  else { ItemGradPol<2> pol{{p}, gbuf, peers}; launch_seg(c, pol, p.mc.DV, false, l); }
}
void launch_label_pass(int NV, const SegCommon& c, const LabelPolParams& p, const Launch& l) {
  // <= ~100 labels, runs of thousands of entries: FR_LABEL_TC chunks per warp-tile
#ifndef FR_LABEL_TC
#define FR_LABEL_TC 8
#endif
  if (p.tab == 1) {
    if (NV == 1) { LabelPol<1, true> pol{p}; launch_seg_tiled<LabelPol<1, true>, FR_LABEL_TC>(c, pol, p.mc.DV, true, l); }
    else { LabelPol<2, true> pol{p}; launch_seg_tiled<LabelPol<2, true>, FR_LABEL_TC>(c, pol, p.mc.DV, true, l); }
    return;
  }
  if (NV == 1) { LabelPol<1> pol{p}; launch_seg_tiled<LabelPol<1>, FR_LABEL_TC>(c, pol, p.mc.DV, true, l); }
  else { LabelPol<2> pol{p}; launch_seg_tiled<LabelPol<2>, FR_LABEL_TC>(c, pol, p.mc.DV, true, l); }
}

// ---- lazy Adam: the batch's unique recipe rows are brought to step-1 before anything
// reads them (forward, dP accumulation, Write_Memory all read R).  keys = recipe ids of the
// item rows, sorted; the warp that sees a run's head owns that row.
template <int NV, int OPT>
__global__ void __launch_bounds__(FR_THREADS)
item_catchup_kernel(const uint32_t* __restrict__ keys, uint32_t n_host, const uint32_t* n_dev,
                    float4* __restrict__ R, float4* __restrict__ m, float4* __restrict__ v,
                    int32_t* __restrict__ last, int DV, const OptConsts oc) {
  const uint32_t n = n_dev ? min(*n_dev, n_host) : n_host;
  const int lane = threadIdx.x & 31;
  const uint32_t gw = blockIdx.x * FR_WARPS_PER_BLOCK + (threadIdx.x >> 5);
  const uint32_t nw = gridDim.x * FR_WARPS_PER_BLOCK;
  const uint32_t nchunks = (n + 31) >> 5;
  const int to = oc.step - 1;
  for (uint32_t chunk = gw; chunk < nchunks; chunk += nw) {
    const uint32_t base = chunk << 5;
    const bool valid = base + lane < n;
    const uint32_t key = valid ? keys[base + lane] : 0xffffffffu;
    const uint32_t prevKey = base > 0 ? keys[base - 1] : 0u;
    const uint32_t up = __shfl_up_sync(FR_FULL, key, 1);
    const bool head = valid && (lane == 0 ? (base == 0 || prevKey != key) : (up != key));
    uint32_t hm = __ballot_sync(FR_FULL, head);
    while (hm) {
      const int e = __ffs(hm) - 1;
      hm &= hm - 1;
      const uint32_t k = __shfl_sync(FR_FULL, key, e);
      const int lastk = last[k];
      __syncwarp();
      if (lastk >= to) continue;
      float4 x[NV], mm[NV], vv[NV];
      load_row<NV>(mm, m + (size_t)k * DV, DV, lane);
      load_row<NV>(vv, v + (size_t)k * DV, DV, lane);
      load_row<NV>(x, R + (size_t)k * DV, DV, lane);
      adam_catchup<OPT, NV>(x, mm, vv, lastk, to, oc);
      store_row<NV>(R + (size_t)k * DV, x, DV, lane);
      store_row<NV>(m + (size_t)k * DV, mm, DV, lane);
      store_row<NV>(v + (size_t)k * DV, vv, DV, lane);
      if (lane == 0) last[k] = to;
    }
  }
}
void launch_item_catchup(int NV, const uint32_t* keys, uint32_t n, float4* R, float4* m, float4* v,
                         int32_t* last, int DV, const OptConsts& oc, const Launch& l, const uint32_t* n_dev) {
  const uint32_t nchunks = (n + 31) / 32;
  if (nchunks == 0) return;
  int grid = (int)((nchunks + FR_WARPS_PER_BLOCK - 1) / FR_WARPS_PER_BLOCK);
  if (grid > l.sm_count * 16) grid = l.sm_count * 16;
  const bool series = oc.adam_mode == FR_ADAM_LAZY_SERIES;
#define FR_ICU(NVV, OO) item_catchup_kernel<NVV, OO><<<grid, FR_THREADS, 0, l.st>>>(keys, n, n_dev, R, m, v, last, DV, oc)
  if (NV == 1) { if (series) FR_ICU(1, OPT_ADAM_SERIES); else FR_ICU(1, OPT_ADAM_EXACT); }
  else         { if (series) FR_ICU(2, OPT_ADAM_SERIES); else FR_ICU(2, OPT_ADAM_EXACT); }
#undef FR_ICU
  ++g_launches;
}

}  // namespace fr

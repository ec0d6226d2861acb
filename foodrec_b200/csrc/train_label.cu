// General_Memory write (Write_Memory, Model_Recommender.py:201-215), every step:
//     G[l] += sum_{rows r : lam_rl != 0}  lam_rl * ws_r * [ beta_2 * pooledCat_r ; beta_1 * m_rc * R[i_r] ]
// i.e. the reference's  tf.matmul(user_label_onehot^T, dish_memory)  with a [L x S] matrix that has 1-3 non-zeros per
// column.  G is tiny (L*5*D floats = 243 KB at L=95, D=128) and the batch is large, so the natural form is a SCATTER
// into an accumulator that lives in shared memory -- not the sort-by-label + segment-reduce the other passes use
// (round 1: label_count/scan/emit + a radix sort of ~1M entries + an issue-bound tile kernel + two combines = 0.25 ms
// of a 1.66 ms step to update 243 KB).
//
// One persistent CTA per SM keeps the slot-1..4 accumulators of its label range [Lp][4][D] in shared memory (194 KB at
// L=95, D=128; wider tables are cut into label ranges, one group of CTAs per range).  Rows are taken in tiles; the
// tile's (row, label, coef) entries are staged in shared memory in (row, label) order.  DETERMINISM WITHOUT ATOMICS:
// every label has exactly one owner warp in the CTA (label % warps); a warp walks the staged entries with ballots,
// picks its own in order and adds into its accumulator rows -- a single writer per accumulator row and a fixed
// order (tile, row, label), so the result is identical run to run.  Slot 0 (beta_2 * pooledCat) is linear in the four
// category rows: only the four weights  sum coef * m_c / n  per label are accumulated; the rows are formed once at
// the end.  Each CTA then writes its accumulator as one partial and label_reduce_kernel adds the partials to G in
// CTA order.
#include "common.cuh"
#include "train.cuh"

namespace fr {

constexpr int LS_THREADS = 512;
constexpr int LS_WARPS = LS_THREADS / 32;
constexpr int LS_TILE = 256;          // rows staged per tile
constexpr int LS_ECAP = 1024;         // entries staged per round (a tile with more entries takes several rounds)

__device__ __forceinline__ int ls_row_labels(const LabelScatterParams& p, int r, int l0, int l1, int first, int limit,
                                             int t, int* e_row, int* e_lab, float* e_coef, float ws) {
  // Walks the labels of item row r that fall in [l0, l1), in ascending label order.  Entries whose running index lies
  // in [first, first + limit) are staged (e_row != nullptr) at index - first.  Returns the number of labels in range.
  const int grp = r / p.group;
  int cnt = 0;
  if (p.user_labels) {
    const float* row = p.user_labels + (size_t)grp * p.mc.L;
    for (int l = l0; l < l1; ++l) {
      const float lam = row[l];
      if (lam != 0.f) {
        if (e_row) {
          const int k = first + cnt;            // (first = this row's offset - window start; may be negative)
          if (k >= 0 && k < limit) { e_row[k] = t; e_lab[k] = l - l0; e_coef[k] = lam * ws; }
        }
        ++cnt;
      }
    }
  } else {
    const int u = p.users[grp];
    const int b = p.lab_off[u], e = p.lab_off[u + 1];
    for (int q = b; q < e; ++q) {
      const int l = p.lab_idx[q];
      if (l >= l0 && l < l1) {
        if (e_row) {
          const int k = first + cnt;
          if (k >= 0 && k < limit) { e_row[k] = t; e_lab[k] = l - l0; e_coef[k] = ws; }
        }
        ++cnt;
      }
    }
  }
  return cnt;
}

template <int NV>
__global__ void __launch_bounds__(LS_THREADS, 1)
label_scatter_kernel(const LabelScatterParams p) {
  extern __shared__ float4 smem[];
  const int DV = p.mc.DV, Lp = p.Lp;
  float4* acc = smem;                                  // [Lp][4][DV]   slots 1..4
  float4* sCat = acc + (size_t)Lp * 4 * DV;            // [4][DV]
  float4* row_mask = sCat + 4 * DV;                    // [LS_TILE]
  float* cw = reinterpret_cast<float*>(row_mask + LS_TILE);   // [Lp][4]  slot-0 category weights
  int* row_item = reinterpret_cast<int*>(cw + Lp * 4); // [LS_TILE]
  float* row_ws = reinterpret_cast<float*>(row_item + LS_TILE);
  int* row_off = reinterpret_cast<int*>(row_ws + LS_TILE);     // [LS_TILE + 1]
  int* wsum = row_off + LS_TILE + 1;                   // [LS_WARPS]
  int* e_row = wsum + LS_WARPS;                        // [LS_ECAP]
  int* e_lab = e_row + LS_ECAP;
  float* e_coef = reinterpret_cast<float*>(e_lab + LS_ECAP);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = blockIdx.x % p.n_parts, cta = blockIdx.x / p.n_parts, ctas = gridDim.x / p.n_parts;
  const int l0 = part * Lp, l1 = min(p.mc.L, l0 + Lp);
  for (int i = tid; i < Lp * 4 * DV; i += LS_THREADS) acc[i] = f4zero();
  for (int i = tid; i < Lp * 4; i += LS_THREADS) cw[i] = 0.f;
  for (int i = tid; i < 4 * DV; i += LS_THREADS) sCat[i] = p.cat[i];
  __syncthreads();

  const int ntiles = (p.S + LS_TILE - 1) / LS_TILE;
  unsigned my_entries = 0;
  for (int tile = cta; tile < ntiles; tile += ctas) {
    const int r = tile * LS_TILE + tid;
    int cnt = 0;
    float ws = 0.f;
    if (tid < LS_TILE) {
      if (r < p.S) {
        const int item = p.items[r];
        ws = p.ws_row[r];
        row_item[tid] = item; row_ws[tid] = ws;
        row_mask[tid] = __ldg(p.cats + (p.cats_by_item ? item : r));
        cnt = ls_row_labels(p, r, l0, l1, 0, 0, tid, nullptr, nullptr, nullptr, ws);
      }
      // exclusive scan of cnt over the tile's rows (8 warps of 32)
      int inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FR_FULL, inc, o); if (lane >= o) inc += v; }
      if (lane == 31) wsum[warp] = inc;
      row_off[tid] = inc - cnt;                        // warp-local exclusive prefix (fixed up below)
    }
    __syncthreads();
    int wbase = 0, tot = 0;
    if (tid < LS_TILE) {
      for (int w = 0; w < LS_TILE / 32; ++w) { const int s = wsum[w]; if (w < warp) wbase += s; tot += s; }
      row_off[tid] += wbase;
    }
    __syncthreads();
    if (tid == 0) { int s = 0; for (int w = 0; w < LS_TILE / 32; ++w) s += wsum[w]; row_off[LS_TILE] = s; }
    __syncthreads();
    tot = row_off[LS_TILE];
    if (tid == 0) my_entries += (unsigned)tot;
    for (int e0 = 0; e0 < tot; e0 += LS_ECAP) {
      if (tid < LS_TILE && r < p.S && cnt > 0) {
        const int off = row_off[tid];
        if (off < e0 + LS_ECAP && off + cnt > e0)
          ls_row_labels(p, r, l0, l1, off - e0, LS_ECAP, tid, e_row, e_lab, e_coef, ws);
      }
      __syncthreads();
      const int n = min(LS_ECAP, tot - e0);
      for (int base = 0; base < n; base += 32) {
        const int idx = base + lane;
        const bool mine = idx < n && (e_lab[idx] % LS_WARPS) == warp;
        uint32_t bal = __ballot_sync(FR_FULL, mine);
        while (bal) {
          // up to 4 of this warp's entries at a time: their recipe rows are requested together
          int ej[4]; int ne = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            ej[q] = -1;
            if (bal) { ej[q] = base + __ffs(bal) - 1; bal &= bal - 1; ++ne; }
          }
          float4 rr[4][NV];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int row = ej[q] >= 0 ? e_row[ej[q]] : e_row[ej[0]];
            load_row_ro<NV>(rr[q], p.R + (size_t)row_item[row] * DV, DV, lane);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (q >= ne) break;
            const int e = ej[q];
            const int row = e_row[e], lab = e_lab[e];
            const float coef = e_coef[e];
            const float4 m = row_mask[row];
            if (lane < 4) {                            // slot 0: category weights  coef * m_c / n   (:124-147)
              const float rn = __frcp_rn(((m.x + m.y) + m.z) + m.w);
              const float mc = comp(m, lane);
              cw[lab * 4 + lane] = __fadd_rn(cw[lab * 4 + lane], __fmul_rn(__fmul_rn(coef, rn), mc));
            }
            const float lc = p.mc.beta_1 * coef;       // slots 1..4: beta_1 * lam * ws * m_c * R[i]   (:111-119)
            float4* a = acc + (size_t)lab * 4 * DV;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float mcv = comp(m, c);
              if (mcv != 0.f) {                        // warp-uniform
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                  const int i = lane + 32 * k;
                  if (i < DV) { float4 v = a[c * DV + i]; mad4_rn(v, lc, scale4(mcv, rr[q][k])); a[c * DV + i] = v; }
                }
              }
            }
          }
        }
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (tid == 0 && my_entries) atomicAdd(p.n_entries, my_entries);      // (integer: order-independent)
  // this CTA's partial: [L][5][DV] rows of its label range
  float4* dst = p.partial + (size_t)cta * p.mc.L * 5 * DV;
  for (int i = tid; i < (l1 - l0) * 4 * DV; i += LS_THREADS) {
    const int l = i / (4 * DV), rem = i % (4 * DV);
    dst[((size_t)(l0 + l) * 5 + 1) * DV + rem] = acc[i];
  }
  for (int i = tid; i < (l1 - l0) * DV; i += LS_THREADS) {
    const int l = i / DV, d = i % DV;
    float4 v = f4zero();
    mad4_rn(v, p.mc.beta_2 * cw[l * 4 + 0], sCat[d]); mad4_rn(v, p.mc.beta_2 * cw[l * 4 + 1], sCat[DV + d]);
    mad4_rn(v, p.mc.beta_2 * cw[l * 4 + 2], sCat[2 * DV + d]); mad4_rn(v, p.mc.beta_2 * cw[l * 4 + 3], sCat[3 * DV + d]);
    dst[((size_t)(l0 + l) * 5) * DV + d] = v;
  }
}

// ---- resident label CSR with <= LS2_MAXL labels per user (the normal case: 1-3): a pipelined variant.
// The general kernel above is latency-bound: staging a tile is a chain of dependent loads (row -> recipe -> mask,
// row -> user -> CSR offsets -> labels) that every warp waits for at a block-wide barrier, and an owner warp finds only
// ~2 of its entries in a 32-entry window, so its recipe-row loads go out two at a time (measured: 0.40 ms per
// 524k-row step, slower than the sort-by-label pass it was meant to replace).  Here
//   * warps 0..7 are STAGERS: they fill one of two staging buffers with tile k+1 (recipe id, sign, mask, the row's
//     labels packed as bytes) while
//   * warps 8..31 are OWNERS (label % 24): an owner first collects ITS (row, label) entries of tile k into a private
//     queue with ballots, then works the queue off eight recipe rows at a time;
//   * one barrier per tile.  Same single-writer / fixed-order rule, so the result is deterministic.
constexpr int LS2_THREADS = 1024;
constexpr int LS2_STAGERS = 8;                       // warps
constexpr int LS2_OWNERS = LS2_THREADS / 32 - LS2_STAGERS;
constexpr int LS2_TILE = LS2_STAGERS * 32;           // rows per tile: one per stager thread
constexpr int LS2_MAXL = 8;                          // labels per row (two packed words)
constexpr int LS2_QCAP = 96;

template <int NV>
__global__ void __launch_bounds__(LS2_THREADS, 1)
label_scatter_csr_kernel(const LabelScatterParams p) {
  extern __shared__ float4 smem[];
  const int DV = p.mc.DV, Lp = p.Lp;
  float4* acc = smem;                                  // [Lp][4][DV]
  float4* sCat = acc + (size_t)Lp * 4 * DV;            // [4][DV]
  float4* row_mask = sCat + 4 * DV;                    // [2][TILE]
  float* cw = reinterpret_cast<float*>(row_mask + 2 * LS2_TILE);   // [Lp][4]
  int* row_item = reinterpret_cast<int*>(cw + Lp * 4); // [2][TILE]
  float* row_ws = reinterpret_cast<float*>(row_item + 2 * LS2_TILE);
  uint32_t* row_labs = reinterpret_cast<uint32_t*>(row_ws + 2 * LS2_TILE);   // [2][TILE][2]: 8 label bytes, 0xff = none
  uint32_t* queue = row_labs + 4 * LS2_TILE;           // [OWNERS][QCAP]: row << 8 | label

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int part = blockIdx.x % p.n_parts, cta = blockIdx.x / p.n_parts, ctas = gridDim.x / p.n_parts;
  const int l0 = part * Lp, l1 = min(p.mc.L, l0 + Lp);
  for (int i = tid; i < Lp * 4 * DV; i += LS2_THREADS) acc[i] = f4zero();
  for (int i = tid; i < Lp * 4; i += LS2_THREADS) cw[i] = 0.f;
  for (int i = tid; i < 4 * DV; i += LS2_THREADS) sCat[i] = p.cat[i];

  const int ntiles = (p.S + LS2_TILE - 1) / LS2_TILE;
  const bool stager = warp < LS2_STAGERS;
  unsigned my_entries = 0;

  auto stage = [&](int tile, int buf) {                // one row per stager thread
    const int r = tile * LS2_TILE + tid;
    uint32_t w0 = 0xffffffffu, w1 = 0xffffffffu;
    if (r < p.S) {
      const int item = p.items[r];
      const int u = p.users[r / p.group];
      const int b = p.lab_off[u], e = p.lab_off[u + 1];
      row_item[buf * LS2_TILE + tid] = item;
      row_ws[buf * LS2_TILE + tid] = p.ws_row[r];
      row_mask[buf * LS2_TILE + tid] = __ldg(p.cats + (p.cats_by_item ? item : r));
      int k = 0;
      for (int q = b; q < e && k < LS2_MAXL; ++q) {
        const int l = p.lab_idx[q];
        if (l >= l0 && l < l1) {
          const uint32_t v = (uint32_t)(l - l0);
          if (k < 4) w0 = (w0 & ~(0xffu << (8 * k))) | (v << (8 * k));
          else w1 = (w1 & ~(0xffu << (8 * (k - 4)))) | (v << (8 * (k - 4)));
          ++k;
        }
      }
      my_entries += (unsigned)k;
    }
    row_labs[(buf * LS2_TILE + tid) * 2] = w0;
    row_labs[(buf * LS2_TILE + tid) * 2 + 1] = w1;
  };

  auto drain = [&](const uint32_t* q, int qn, int buf) {     // this owner's entries, eight recipe rows in flight
    constexpr int PF = NV == 1 ? 8 : 4;
    for (int i0 = 0; i0 < qn; i0 += PF) {
      uint32_t ent[PF]; float4 rr[PF][NV];
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        ent[k] = q[min(i0 + k, qn - 1)];
        load_row_ro<NV>(rr[k], p.R + (size_t)row_item[buf * LS2_TILE + (ent[k] >> 8)] * DV, DV, lane);
      }
#pragma unroll
      for (int k = 0; k < PF; ++k) {
        if (i0 + k >= qn) break;
        const int row = (int)(ent[k] >> 8), lab = (int)(ent[k] & 0xffu);
        const float coef = row_ws[buf * LS2_TILE + row];
        const float4 m = row_mask[buf * LS2_TILE + row];
        if (lane < 4) {
          const float rn = __frcp_rn(((m.x + m.y) + m.z) + m.w);
          cw[lab * 4 + lane] = __fadd_rn(cw[lab * 4 + lane], __fmul_rn(__fmul_rn(coef, rn), comp(m, lane)));
        }
        const float lc = p.mc.beta_1 * coef;
        float4* a = acc + (size_t)lab * 4 * DV;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float mcv = comp(m, c);
          if (mcv != 0.f) {
#pragma unroll
            for (int x = 0; x < NV; ++x) {
              const int i = lane + 32 * x;
              if (i < DV) { float4 v = a[c * DV + i]; mad4_rn(v, lc, scale4(mcv, rr[k][x])); a[c * DV + i] = v; }
            }
          }
        }
      }
    }
  };

  int tile = cta;
  if (stager && tile < ntiles) stage(tile, 0);
  __syncthreads();
  for (int k = 0; tile < ntiles; ++k, tile += ctas) {
    const int buf = k & 1;
    if (stager) {
      if (tile + ctas < ntiles) stage(tile + ctas, buf ^ 1);
    } else {
      const int o = warp - LS2_STAGERS;
      uint32_t* q = queue + o * LS2_QCAP;
      int qn = 0;
      const int nrows = min(LS2_TILE, p.S - tile * LS2_TILE);
      for (int base = 0; base < nrows; base += 32) {
        const int row = base + lane;
        const uint32_t w0 = row < nrows ? row_labs[(buf * LS2_TILE + row) * 2] : 0xffffffffu;
        const uint32_t w1 = row < nrows ? row_labs[(buf * LS2_TILE + row) * 2 + 1] : 0xffffffffu;
        if (__all_sync(FR_FULL, w1 == 0xffffffffu && (w0 >> 24) == 0xffu)) {   // <= 3 labels everywhere (the usual case)
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const uint32_t lab = (w0 >> (8 * s)) & 0xffu;
            const bool mine = lab != 0xffu && (int)(lab % LS2_OWNERS) == o;
            const uint32_t bal = __ballot_sync(FR_FULL, mine);
            if (mine) q[qn + __popc(bal & ((1u << lane) - 1u))] = ((uint32_t)row << 8) | lab;
            qn += __popc(bal);
            if (qn > LS2_QCAP - 32) { __syncwarp(); drain(q, qn, buf); __syncwarp(); qn = 0; }
          }
        } else {
#pragma unroll
          for (int s = 0; s < LS2_MAXL; ++s) {
            const uint32_t lab = ((s < 4 ? w0 : w1) >> (8 * (s & 3))) & 0xffu;
            const bool mine = lab != 0xffu && (int)(lab % LS2_OWNERS) == o;
            const uint32_t bal = __ballot_sync(FR_FULL, mine);
            if (mine) q[qn + __popc(bal & ((1u << lane) - 1u))] = ((uint32_t)row << 8) | lab;
            qn += __popc(bal);
            if (qn > LS2_QCAP - 32) { __syncwarp(); drain(q, qn, buf); __syncwarp(); qn = 0; }
          }
        }
      }
      __syncwarp();
      drain(q, qn, buf);
    }
    __syncthreads();
  }
  if (stager) {
    my_entries = __reduce_add_sync(FR_FULL, my_entries);
    if (lane == 0 && my_entries) atomicAdd(p.n_entries, my_entries);
  }
  float4* dst = p.partial + (size_t)cta * p.mc.L * 5 * DV;
  for (int i = tid; i < (l1 - l0) * 4 * DV; i += LS2_THREADS) {
    const int l = i / (4 * DV), rem = i % (4 * DV);
    dst[((size_t)(l0 + l) * 5 + 1) * DV + rem] = acc[i];
  }
  for (int i = tid; i < (l1 - l0) * DV; i += LS2_THREADS) {
    const int l = i / DV, d = i % DV;
    float4 v = f4zero();
    mad4_rn(v, p.mc.beta_2 * cw[l * 4 + 0], sCat[d]); mad4_rn(v, p.mc.beta_2 * cw[l * 4 + 1], sCat[DV + d]);
    mad4_rn(v, p.mc.beta_2 * cw[l * 4 + 2], sCat[2 * DV + d]); mad4_rn(v, p.mc.beta_2 * cw[l * 4 + 3], sCat[3 * DV + d]);
    dst[((size_t)(l0 + l) * 5) * DV + d] = v;
  }
}

// G += partials, summed in CTA order (fixed tree: deterministic)
__global__ void __launch_bounds__(256)
label_reduce_kernel(float4* __restrict__ G, const float4* __restrict__ partial, int n4, int n_partials,
                    const uint32_t* __restrict__ n_entries, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && out) out[FR_OUT_LABEL_ENTRIES] = (float)*n_entries;
  if (i >= n4) return;
  float4 s = f4zero();
  int c = 0;
  for (; c + 4 <= n_partials; c += 4) {
    const float4 a = __ldcg(partial + (size_t)c * n4 + i), b = __ldcg(partial + (size_t)(c + 1) * n4 + i);
    const float4 d = __ldcg(partial + (size_t)(c + 2) * n4 + i), e = __ldcg(partial + (size_t)(c + 3) * n4 + i);
    s = add4(add4(add4(add4(s, a), b), d), e);
  }
  for (; c < n_partials; ++c) s = add4(s, __ldcg(partial + (size_t)c * n4 + i));
  G[i] = add4(G[i], s);
}

size_t label_scatter_smem(int Lp, int DV) {
  return ((size_t)Lp * 4 * DV + 4 * DV + LS_TILE) * sizeof(float4) +
         ((size_t)Lp * 4 + 3 * LS_TILE + 1 + LS_WARPS + 3 * LS_ECAP) * 4;
}
size_t label_scatter_csr_smem(int Lp, int DV) {
  return ((size_t)Lp * 4 * DV + 4 * DV + 2 * LS2_TILE) * sizeof(float4) +
         ((size_t)Lp * 4 + 2 * LS2_TILE * 2 + 4 * LS2_TILE + (size_t)LS2_OWNERS * LS2_QCAP) * 4;
}

// Returns false when the accumulator does not fit shared memory even in 8 label ranges (caller takes the sort path).
// csr: the pipelined kernel (resident label CSR, <= LS2_MAXL labels per user, label ranges of <= 255 labels).
bool label_scatter_plan(int L, int DV, int sm_count, bool csr, int* n_parts, int* Lp) {
  static int max_smem = -1;
  if (max_smem < 0) {
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaFuncSetAttribute(label_scatter_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(label_scatter_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(label_scatter_csr_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
    cudaFuncSetAttribute(label_scatter_csr_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  }
  for (int np = 1; np <= 8 && np <= sm_count; ++np) {
    const int lp = (L + np - 1) / np;
    if (csr && lp > 255) continue;
    const size_t need = csr ? label_scatter_csr_smem(lp, DV) : label_scatter_smem(lp, DV);
    if (need <= (size_t)max_smem) { *n_parts = np; *Lp = lp; return true; }
  }
  return false;
}

int label_scatter_ctas(int S, int n_parts, int sm_count, int tile) {       // CTAs per label range (= partials to reduce)
  const int ntiles = (S + tile - 1) / tile;
  int c = sm_count / n_parts;
  if (c > ntiles) c = ntiles;
  return c < 1 ? 1 : c;
}

void launch_label_scatter(int NV, LabelScatterParams p, bool csr, const Launch& l) {
  const int ctas = label_scatter_ctas(p.S, p.n_parts, l.sm_count, csr ? LS2_TILE : LS_TILE);
  cudaMemsetAsync(p.n_entries, 0, sizeof(uint32_t), l.st);
  if (csr) {
    const size_t smem = label_scatter_csr_smem(p.Lp, p.mc.DV);
    if (NV == 1) label_scatter_csr_kernel<1><<<ctas * p.n_parts, LS2_THREADS, smem, l.st>>>(p);
    else label_scatter_csr_kernel<2><<<ctas * p.n_parts, LS2_THREADS, smem, l.st>>>(p);
  } else {
    const size_t smem = label_scatter_smem(p.Lp, p.mc.DV);
    if (NV == 1) label_scatter_kernel<1><<<ctas * p.n_parts, LS_THREADS, smem, l.st>>>(p);
    else label_scatter_kernel<2><<<ctas * p.n_parts, LS_THREADS, smem, l.st>>>(p);
  }
  const int n4 = p.mc.L * 5 * p.mc.DV;
  label_reduce_kernel<<<(n4 + 255) / 256, 256, 0, l.st>>>(p.G, p.partial, n4, ctas, p.n_entries, p.out);
  g_launches += 2;
}

// largest number of labels any user has in the resident CSR (fr_set_tables; synchronises)
__global__ void csr_max_count_kernel(const int32_t* __restrict__ off, int64_t n, int32_t* __restrict__ out) {
  int m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, off[i + 1] - off[i]);
  m = __reduce_max_sync(FR_FULL, m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}
int csr_max_count(const int32_t* off, int64_t n, int32_t* scratch, cudaStream_t st) {
  if (n <= 0) return 0;
  cudaMemsetAsync(scratch, 0, sizeof(int32_t), st);
  csr_max_count_kernel<<<(int)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256), 256, 0, st>>>(off, n, scratch);
  int v = 0;
  cudaMemcpyAsync(&v, scratch, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  return v;
}

}  // namespace fr

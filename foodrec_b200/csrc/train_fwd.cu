// Train-step kernels that are not segment reductions (sm_100a):
//   prep_rows_kernel    row keys + write sign
//   fwd_train_kernel    Model_Recommender.py:56-104 + per-slice gradient norms (:236-237)
//   gcat_reduce/finalize clip_by_global_norm scale (:237) + dense optimizer on Category_Embedding
//   label_count/emit    non-zeros of the label feed -> (label, row, coef) entries
//   adam_sweep_kernel   TF-1.x non-lazy Adam decay of untouched rows (dense sweep / lazy flush)
//   series_*_kernel     LAZY_SERIES coefficient table
#include "optim.cuh"
#include "train.cuh"

namespace fr {

// ------------------------------------------------------------------ prep
// Also the id range check of the step (the reference's tf.gather raises on the CPU for an id outside its table,
// Model_Recommender.py:57,63): every later kernel reads users_s / items_s, in which an out-of-range id is replaced
// by row 0 and FR_OUT_OVERFLOW is set to 3 -- nothing is ever read or written outside a table, and the step's
// result is reported as invalid (Engine.read_scalars raises) instead of silently corrupting rows.
__global__ void prep_rows_kernel(int mode, int B, const int32_t* __restrict__ users, const int32_t* __restrict__ items,
                                 const float* __restrict__ labels, const float* __restrict__ ws_in,
                                 uint32_t n_users, uint32_t n_items,
                                 uint32_t* __restrict__ ukeys, float* __restrict__ ws_row,
                                 int32_t* __restrict__ users_s, int32_t* __restrict__ items_s, float* __restrict__ flag) {
  const int group = mode == FR_BPR ? 2 : 1;
  const int S = B * group;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < S; r += gridDim.x * blockDim.x) {
    const int grp = r / group;
    uint32_t u = (uint32_t)users[grp], it = (uint32_t)items[r];
    if (u >= n_users || it >= n_items) {
      *flag = 3.f;
      if (u >= n_users) u = 0;
      if (it >= n_items) it = 0;
    }
    ukeys[r] = u;
    if (r == grp * group) users_s[grp] = (int32_t)u;
    items_s[r] = (int32_t)it;
    float w;
    if (ws_in) w = ws_in[r];
    else if (mode == FR_BPR) w = (r & 1) ? -1.f : 1.f;
    else w = labels[grp] > 0.5f ? 1.f : -1.f;
    ws_row[r] = w;
  }
}

void launch_prep_rows(int mode, int B, const int32_t* users, const int32_t* items, const float* labels, const float* ws_in,
                      int64_t n_users, int64_t n_items, uint32_t* ukeys, float* ws_row, int32_t* users_s, int32_t* items_s,
                      float* flag, const Launch& l) {
  const int S = B * (mode == FR_BPR ? 2 : 1);
  int grid = (S + 255) / 256;
  if (grid > l.sm_count * 8) grid = l.sm_count * 8;
  if (grid < 1) grid = 1;
  prep_rows_kernel<<<grid, 256, 0, l.st>>>(mode, B, users, items, labels, ws_in, (uint32_t)n_users, (uint32_t)n_items, ukeys,
                                           ws_row, users_s, items_s, flag);
  ++g_launches;
}

// ------------------------------------------------------------------ forward + loss + norms
// One warp per group (sample / BPR triple).  Reads 5D floats of P[u] and D floats of
// R[i] per item row (24*D+32 bytes per sample), writes the z-row stash
// (1-a)/n * sum_c m_c P[u,1+c] so the recipe-gradient pass never re-reads P.
// LAZY = 0 (rows in memory are current) | OPT_ADAM_EXACT | OPT_ADAM_SERIES: P[u] in memory may
// be stale and is brought to step-1 in registers from (m, v) before it is scored.
// (__launch_bounds__(FR_THREADS, 3) -> 80 registers, 24 warps/SM: measured SLOWER, 0.78 vs 0.65 ms, spills)
// TAB = storage format of the tables (fr_set_table_format): 0 fp32 P and R; 1 bf16 P and R; 2 bf16 P, fp32 R (the
// row-sharded step: R = the receive buffer, filled in fp32 by the owners' gather).  bf16 only without lazy Adam.
template <int NV, int GROUP, int LAZY, int TAB>
__global__ void __launch_bounds__(FR_THREADS)
fwd_train_kernel(const FwdParams p) {
  constexpr bool BFP = TAB != 0, BFR = TAB == 1;
  static_assert(TAB == 0 || LAZY == 0, "bf16 tables: no lazy Adam");
  extern __shared__ float4 smem[];
  const int DV = p.DV;
  float4* sCat = smem;                 // [4*DV]
  float4* red = smem + 4 * DV;         // [WARPS][4*DV]
  __shared__ float red_loss[FR_WARPS_PER_BLOCK], red_nrm[FR_WARPS_PER_BLOCK];
  for (int i = threadIdx.x; i < 4 * DV; i += blockDim.x) sCat[i] = p.cat[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * FR_WARPS_PER_BLOCK + warp, nw = gridDim.x * FR_WARPS_PER_BLOCK;
  const float a = p.a, oma = p.oma, Bf = p.Bnorm;
  float4 gc[4][NV];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < NV; ++k) gc[c][k] = f4zero();
  float lossacc = 0.f, nrmacc = 0.f;

  // Software pipeline over this warp's groups.  The chain  id -> stamp / category mask -> rows  is three dependent
  // loads; the profile (profiles/r01_ncu_full_train_B262144.csv, source page) showed the warp parked on each of
  // them.  Ids are therefore loaded TWO groups ahead and the stamp, the masks and an L2 prefetch of the rows ONE
  // group ahead, so that the current group only waits for its row loads.
  int u = 0, lastu = 0, it[GROUP]; float4 mc[GROUP];       // current group
  int u1 = 0, it1[GROUP];                                   // ids of the next group
#pragma unroll
  for (int j = 0; j < GROUP; ++j) { it[j] = 0; it1[j] = 0; mc[j] = make_float4(1.f, 0.f, 0.f, 0.f); }
  if (gw < p.B) {
    u = p.users[gw];
    if constexpr (LAZY != 0) lastu = p.lastP[u];
#pragma unroll
    for (int j = 0; j < GROUP; ++j) {
      it[j] = p.items[gw * GROUP + j];
      mc[j] = __ldg(p.cats + (p.cats_by_item ? it[j] : gw * GROUP + j));
    }
    if (gw + nw < p.B) {
      u1 = p.users[gw + nw];
#pragma unroll
      for (int j = 0; j < GROUP; ++j) it1[j] = p.items[(gw + nw) * GROUP + j];
    }
  }
  for (int grp = gw; grp < p.B; grp += nw) {
    const int gn = grp + nw, gnn = grp + 2 * nw;
    int u2 = 0, it2[GROUP], last1 = 0; float4 mc1[GROUP];
#pragma unroll
    for (int j = 0; j < GROUP; ++j) { it2[j] = 0; mc1[j] = make_float4(1.f, 0.f, 0.f, 0.f); }
    if (gnn < p.B) {            // ids two groups ahead: nothing below depends on them
      u2 = p.users[gnn];
#pragma unroll
      for (int j = 0; j < GROUP; ++j) it2[j] = p.items[gnn * GROUP + j];
    }
    if (gn < p.B) {             // next group: stamp, masks, and its rows into L2 (no registers held)
      if constexpr (LAZY != 0) last1 = p.lastP[u1];
#pragma unroll
      for (int j = 0; j < GROUP; ++j) mc1[j] = __ldg(p.cats + (p.cats_by_item ? it1[j] : gn * GROUP + j));
#ifndef FR_NO_FWD_PREFETCH
      const uint32_t ub = 5u * (uint32_t)DV * (BFP ? 8u : 16u), rb = (uint32_t)DV * (BFR ? 8u : 16u);
      const size_t uo = (size_t)u1 * 5 * DV;
      prefetch_l2_warp(tab_at<BFP>(p.P, uo), ub, lane);
      if (LAZY != 0) { prefetch_l2_warp(p.mP + uo, ub, (lane + 31) & 31); prefetch_l2_warp(p.vP + uo, ub, (lane + 30) & 31); }
#pragma unroll
      for (int j = 0; j < GROUP; ++j)
        prefetch_l2_warp(tab_at<BFR>(p.R, (size_t)it1[j] * DV), rb, (lane + 29 - j) & 31);
#endif
    }
    float4 pr[5][NV];
#pragma unroll
    for (int s = 0; s < 5; ++s) load_row_t<NV>(pr[s], tab_at<BFP>(p.P, ((size_t)u * 5 + s) * DV), DV, lane);
    float4 rr[GROUP][NV], pcn[GROUP][NV];   // R rows, pooledCat (normalised)
#pragma unroll
    for (int j = 0; j < GROUP; ++j) load_row_ro_t<NV>(rr[j], tab_at<BFR>(p.R, (size_t)it[j] * DV), DV, lane);
    if constexpr (LAZY != 0) {
      const int to = p.oc.step - 1;
      if (lastu < to) {
        float4 mm[5][NV], vv[5][NV];
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          load_row<NV>(mm[s], p.mP + ((size_t)u * 5 + s) * DV, DV, lane);
          load_row<NV>(vv[s], p.vP + ((size_t)u * 5 + s) * DV, DV, lane);
        }
        adam_catchup<LAZY, 5 * NV>(&pr[0][0], &mm[0][0], &vv[0][0], lastu, to, p.oc);
      }
    }
    float4 mm[GROUP];
    // hs, ls are reduced across the warp (the score needs them); the squared norms stay per-lane partials and are
    // reduced once, at the end of the kernel
    float sc[GROUP], rnn[GROUP], nzq[GROUP], nRq[GROUP], npcq[GROUP];
#pragma unroll
    for (int j = 0; j < GROUP; ++j) {
      const int r = grp * GROUP + j;
      const float4 m = mc[j];
      const float n = ((m.x + m.y) + m.z) + m.w;                    // :77
      // one IEEE reciprocal per item row; x * (1/n) is x / n exactly for n = 1, 2, 4 and within 1 ulp for n = 3
      // (same rule as row_terms in train_seg.cu) -- eleven divisions less per row
      const float rn = __frcp_rn(n);
      float4 pcs[NV], zs[NV];
      pooled_cat<NV>(pcs, sCat, m, DV, lane);
      float hs = 0.f, ls = 0.f, nz = 0.f, nR = 0.f, npc = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        zs[k].x = m.x * pr[1][k].x + m.y * pr[2][k].x + m.z * pr[3][k].x + m.w * pr[4][k].x;
        zs[k].y = m.x * pr[1][k].y + m.y * pr[2][k].y + m.z * pr[3][k].y + m.w * pr[4][k].y;
        zs[k].z = m.x * pr[1][k].z + m.y * pr[2][k].z + m.z * pr[3][k].z + m.w * pr[4][k].z;
        zs[k].w = m.x * pr[1][k].w + m.y * pr[2][k].w + m.z * pr[3][k].w + m.w * pr[4][k].w;
        hs += dot4(pr[0][k], pcs[k]);
        ls += dot4(zs[k], rr[j][k]);
        nz += dot4(zs[k], zs[k]);
        if constexpr (GROUP == 1) { nR += dot4(rr[j][k], rr[j][k]); npc += dot4(pcs[k], pcs[k]); }
      }
      hs = warp_sum(hs); ls = warp_sum(ls);
      const float high = hs * rn, low = ls * rn;                    // :79, :92
      sc[j] = a * high + oma * low;                                 // :95-96
      const float zc = oma * rn;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        if (i < DV) __stcg(p.z + (size_t)r * DV + i, scale4(zc, zs[k]));
        pcn[j][k] = scale4(rn, pcs[k]);
      }
      mm[j] = m; rnn[j] = rn;
      const float inv2 = rn * rn;
      nzq[j] = nz * inv2; nRq[j] = nR * inv2; npcq[j] = npc * inv2;
    }
    if constexpr (GROUP == 1) {
      const float s = sc[0], y = p.labels[grp];
      const float e = expf(-fabsf(s));
      const float loss = fmaxf(s, 0.f) - s * y + log1pf(e);         // :101
      const float sig = s >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
      const float g = (sig - y) / Bf;
      const float4 m = mm[0];
      const float sumsq_m = m.x * m.x + m.y * m.y + m.z * m.z + m.w * m.w;
      const float nrm = g * g * (a * a * npcq[0] + oma * oma * (nRq[0] * sumsq_m + nzq[0]));
      const float ga = g * a, rn = rnn[0];
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        fma4(gc[0][k], ga * (m.x * rn), pr[0][k]); fma4(gc[1][k], ga * (m.y * rn), pr[0][k]);
        fma4(gc[2][k], ga * (m.z * rn), pr[0][k]); fma4(gc[3][k], ga * (m.w * rn), pr[0][k]);
      }
      lossacc += loss; nrmacc += nrm;
      if (lane == 0) { p.g[grp] = g; p.scores[grp] = s; }
    } else {
      const float s = sc[0] - sc[GROUP - 1];
      const float e = expf(-fabsf(s));
      const float loss = fmaxf(s, 0.f) - s + log1pf(e);
      const float sig = s >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
      const float h = (sig - 1.f) / Bf;
      const float4 m0 = mm[0], m1 = mm[GROUP - 1];
      const float4 w0 = scale4(rnn[0], m0), w1 = scale4(rnn[GROUP - 1], m1);
      // || h (q0 - q1) ||^2 of the P slice of this triple
      float dq = 0.f;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const float4 r0 = rr[0][k], r1 = rr[GROUP - 1][k];
        float4 d = make_float4(pcn[0][k].x - pcn[GROUP - 1][k].x, pcn[0][k].y - pcn[GROUP - 1][k].y,
                               pcn[0][k].z - pcn[GROUP - 1][k].z, pcn[0][k].w - pcn[GROUP - 1][k].w);
        dq += a * a * dot4(d, d);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float wa = comp(w0, c), wb = comp(w1, c);
          d = make_float4(wa * r0.x - wb * r1.x, wa * r0.y - wb * r1.y, wa * r0.z - wb * r1.z, wa * r0.w - wb * r1.w);
          dq += oma * oma * dot4(d, d);
        }
      }
      const float nrm = h * h * (dq + oma * oma * (nzq[0] + nzq[GROUP - 1]));     // per-lane partial
      const float ha = h * a;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        fma4(gc[0][k], ha * (w0.x - w1.x), pr[0][k]); fma4(gc[1][k], ha * (w0.y - w1.y), pr[0][k]);
        fma4(gc[2][k], ha * (w0.z - w1.z), pr[0][k]); fma4(gc[3][k], ha * (w0.w - w1.w), pr[0][k]);
      }
      lossacc += loss; nrmacc += nrm;
      if (lane == 0) {
        p.g[grp * GROUP] = h; p.g[grp * GROUP + 1] = -h;
        p.scores[grp * GROUP] = sc[0]; p.scores[grp * GROUP + 1] = sc[GROUP - 1];
      }
    }
      u = u1; lastu = last1; u1 = u2;
#pragma unroll
    for (int j = 0; j < GROUP; ++j) { it[j] = it1[j]; mc[j] = mc1[j]; it1[j] = it2[j]; }
  }
  // deterministic block reduction (fixed warp order), one partial per block
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      if (i < DV) red[(warp * 4 + c) * DV + i] = gc[c][k];
    }
  nrmacc = warp_sum(nrmacc);
  if (lane == 0) { red_loss[warp] = lossacc; red_nrm[warp] = nrmacc; }
  __syncthreads();
  for (int j = threadIdx.x; j < 4 * DV; j += blockDim.x) {
    float4 s = red[j];
#pragma unroll
    for (int w = 1; w < FR_WARPS_PER_BLOCK; ++w) s = add4(s, red[w * 4 * DV + j]);
    p.part_gcat[(size_t)blockIdx.x * 4 * DV + j] = s;
  }
  if (threadIdx.x == 0) {
    float l = 0.f, q = 0.f;
#pragma unroll
    for (int w = 0; w < FR_WARPS_PER_BLOCK; ++w) { l += red_loss[w]; q += red_nrm[w]; }
    p.part_loss[blockIdx.x] = l; p.part_nrm[blockIdx.x] = q;
  }
}

int fwd_train_grid(int B, int sm_count) {
  int grid = (B + FR_WARPS_PER_BLOCK - 1) / FR_WARPS_PER_BLOCK;
  // 4 CTAs per SM although 3 are resident: measured faster (the block scheduler evens out
  // the per-warp work); one dCat partial per CTA (workspace is sized for sm_count*4)
  const int cap = sm_count * 4;
  if (grid > cap) grid = cap;
  return grid < 1 ? 1 : grid;
}

void launch_fwd_train(int NV, int group, const FwdParams& p, int grid, const Launch& l) {
  const size_t smem = (size_t)(4 * p.DV) * sizeof(float4) * (1 + FR_WARPS_PER_BLOCK);
  const int lz = !p.lazy ? 0 : (p.oc.adam_mode == FR_ADAM_LAZY_SERIES ? OPT_ADAM_SERIES : OPT_ADAM_EXACT);
#define FR_FWD(NVV, GG, LZ, TB) fwd_train_kernel<NVV, GG, LZ, TB><<<grid, FR_THREADS, smem, l.st>>>(p)
#define FR_FWD_L(NVV, GG) do { if (p.tab == 1) FR_FWD(NVV, GG, 0, 1); else if (p.tab == 2) FR_FWD(NVV, GG, 0, 2);             \
                               else if (lz == 0) FR_FWD(NVV, GG, 0, 0); else if (lz == OPT_ADAM_EXACT) FR_FWD(NVV, GG, OPT_ADAM_EXACT, 0); \
                               else FR_FWD(NVV, GG, OPT_ADAM_SERIES, 0); } while (0)
  if (NV == 1) { if (group == 1) FR_FWD_L(1, 1); else FR_FWD_L(1, 2); }
  else         { if (group == 1) FR_FWD_L(2, 1); else FR_FWD_L(2, 2); }
#undef FR_FWD_L
#undef FR_FWD
  ++g_launches;
}

// ------------------------------------------------------------------ finalize
// Single block.  reduce: block partials -> packed {loss_sum, nrm_sum, gCat}.
// apply: global norm, clip scale (clip_ops.py: clip * min(1/norm, 1/clip)), scalars,
// dense optimizer on Category_Embedding (ApplyAdam / ApplyAdagrad / ApplyRMSProp / SGD).
__global__ void __launch_bounds__(FR_THREADS)
finalize_kernel(const FinalizeParams p) {
  __shared__ double sh[FR_WARPS_PER_BLOCK];
  __shared__ float s_scale;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n4 = 4 * p.DV;
  float4* pk_gcat = reinterpret_cast<float4*>(p.packed + 4);   // packed[0..3] = loss, nrm, pad, pad
  if (p.do_reduce) {      // (dCat partials were reduced by gcat_reduce_kernel just before)
    if (warp == 0) {
      double l = 0.0, q = 0.0;
      for (int b = lane; b < p.nblk; b += 32) { l += (double)p.part_loss[b]; q += (double)p.part_nrm[b]; }
      l = warp_sum_d(l); q = warp_sum_d(q);
      if (lane == 0) { p.packed[0] = (float)l; p.packed[1] = (float)q; }
    }
    __syncthreads();
  }
  if (!p.do_apply) return;
  double sq = 0.0;
  for (int j = threadIdx.x; j < n4; j += blockDim.x) {
    const float4 g = pk_gcat[j];
    sq += (double)g.x * g.x + (double)g.y * g.y + (double)g.z * g.z + (double)g.w * g.w;
  }
  sq = warp_sum_d(sq);
  if (lane == 0) sh[warp] = sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < FR_WARPS_PER_BLOCK; ++w) t += sh[w];
    const float norm = sqrtf((float)(t + (double)p.packed[1]));
    const float scale = p.clip * fminf(1.f / norm, 1.f / p.clip);
    s_scale = scale;
    p.out[FR_OUT_LOSS] = p.packed[0] / p.B;
    p.out[FR_OUT_NORM] = norm;
    p.out[FR_OUT_SCALE] = scale;
    p.out[FR_OUT_LR] = p.oc.lr;
    if (p.lr_hist) p.lr_hist[p.oc.step] = p.oc.lr_t;
  }
  __syncthreads();
  const float scale = s_scale;
  const OptConsts& oc = p.oc;
  float* var = reinterpret_cast<float*>(p.Cat);
  float* s1 = reinterpret_cast<float*>(p.s1Cat);
  float* s2 = reinterpret_cast<float*>(p.s2Cat);
  const float* gp = reinterpret_cast<const float*>(pk_gcat);
  for (int j = threadIdx.x; j < 4 * n4; j += blockDim.x) {
    const float g = gp[j] * scale;
    float x = var[j];
    if (oc.learner == FR_ADAM) {         // training_ops.cc ApplyAdam (dense form)
      float m = s1[j], v = s2[j];
      m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), oc.omb1));
      v = __fadd_rn(v, __fmul_rn(__fsub_rn(__fmul_rn(g, g), v), oc.omb2));
      x = __fsub_rn(x, __fdiv_rn(__fmul_rn(m, oc.lr_t), __fadd_rn(__fsqrt_rn(v), oc.eps)));
      s1[j] = m; s2[j] = v;
    } else if (oc.learner == FR_ADAGRAD) {
      float acc = s1[j]; adagrad_touch<false>(x, acc, g, oc); s1[j] = acc;
    } else if (oc.learner == FR_RMSPROP) {
      float ms = s1[j], mom = s2[j]; rmsprop_touch<false>(x, ms, mom, g, oc); s1[j] = ms; s2[j] = mom;
    } else {
      sgd_touch(x, g, oc);
    }
    var[j] = x;
  }
}

// dCat block partials [nblk][4*DV] -> [4*DV]: 8 float4 columns per block, 32 row groups per
// column, fixed-order shared-memory tree (deterministic).
__global__ void __launch_bounds__(256)
gcat_reduce_kernel(const float4* __restrict__ part, int nblk, int n4, float4* __restrict__ out) {
  __shared__ float4 sh[32][8];
  const int c = threadIdx.x & 7, r = threadIdx.x >> 3;
  const int j = blockIdx.x * 8 + c;
  float4 s = f4zero();
  if (j < n4)
    for (int b = r; b < nblk; b += 32) s = add4(s, part[(size_t)b * n4 + j]);
  sh[r][c] = s;
  __syncthreads();
  if (r == 0 && j < n4) {
    float4 t = sh[0][c];
#pragma unroll
    for (int q = 1; q < 32; ++q) t = add4(t, sh[q][c]);
    out[j] = t;
  }
}

void launch_finalize(const FinalizeParams& p, const Launch& l) {
  if (p.do_reduce) {
    const int n4 = 4 * p.DV;
    gcat_reduce_kernel<<<(n4 + 7) / 8, 256, 0, l.st>>>(p.part_gcat, p.nblk, n4, reinterpret_cast<float4*>(p.packed + 4));
    ++g_launches;
  }
  finalize_kernel<<<1, FR_THREADS, 0, l.st>>>(p);
  ++g_launches;
}

// ------------------------------------------------------------------ label feed -> entries
// One warp per item row r: the non-zeros (l, lam) of its user's label row become entries
// (key=l, row=r, coef=lam*ws_r) in (r, l) order, so the stable sort by l keeps batch order.
__device__ __forceinline__ int label_row_count(const LabelEmitParams& p, int r, int lane) {
  const int grp = r / p.group;
  if (p.user_labels) {
    int cnt = 0;
    for (int l0 = 0; l0 < p.L; l0 += 32) {
      const int l = l0 + lane;
      const bool nz = l < p.L && p.user_labels[(size_t)grp * p.L + l] != 0.f;
      cnt += __popc(__ballot_sync(FR_FULL, nz));
    }
    return cnt;
  }
  const int u = p.users[grp];
  return p.lab_off[u + 1] - p.lab_off[u];
}

__global__ void __launch_bounds__(FR_THREADS)
label_count_kernel(const LabelEmitParams p) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * FR_WARPS_PER_BLOCK + (threadIdx.x >> 5), nw = gridDim.x * FR_WARPS_PER_BLOCK;
  for (int r = gw; r < p.S; r += nw) {
    const int c = label_row_count(p, r, lane);
    if (lane == 0) p.counts[r] = (uint32_t)c;
  }
}

__global__ void __launch_bounds__(FR_THREADS)
label_emit_kernel(const LabelEmitParams p) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * FR_WARPS_PER_BLOCK + (threadIdx.x >> 5), nw = gridDim.x * FR_WARPS_PER_BLOCK;
  if (gw == 0 && lane == 0) {
    const uint32_t tot = *p.n_entries;
    p.out[FR_OUT_LABEL_ENTRIES] = (float)tot;
    if (tot > p.cap) p.out[FR_OUT_OVERFLOW] = 1.f;      // (slot is zeroed at step start)
  }
  for (int r = gw; r < p.S; r += nw) {
    const int grp = r / p.group;
    const float ws = p.ws_row[r];
    uint32_t off = p.offs[r];
    if (p.user_labels) {
      for (int l0 = 0; l0 < p.L; l0 += 32) {
        const int l = l0 + lane;
        const float lam = l < p.L ? p.user_labels[(size_t)grp * p.L + l] : 0.f;
        const uint32_t bal = __ballot_sync(FR_FULL, lam != 0.f);
        if (lam != 0.f) {
          const uint32_t dst = off + __popc(bal & ((1u << lane) - 1u));
          if (dst < p.cap) { p.ent_key[dst] = (uint32_t)l; p.ent_row[dst] = (uint32_t)r; p.ent_coef[dst] = lam * ws; }
        }
        off += __popc(bal);
      }
    } else {
      const int u = p.users[grp];
      const int b = p.lab_off[u], cnt = p.lab_off[u + 1] - b;
      for (int q = lane; q < cnt; q += 32) {
        const uint32_t dst = off + q;
        if (dst < p.cap) { p.ent_key[dst] = (uint32_t)p.lab_idx[b + q]; p.ent_row[dst] = (uint32_t)r; p.ent_coef[dst] = ws; }
      }
    }
  }
}

// Resident label CSR (ids-only feed): a row has 1-3 labels, so one THREAD per item row (a warp per row spends
// three dependent loads of latency on <= 3 useful lanes: 2 x 100 us per 524k rows, measured).
__global__ void __launch_bounds__(256)
label_count_csr_kernel(const LabelEmitParams p) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < p.S; r += gridDim.x * blockDim.x) {
    const int u = p.users[r / p.group];
    p.counts[r] = (uint32_t)(p.lab_off[u + 1] - p.lab_off[u]);
  }
}
__global__ void __launch_bounds__(256)
label_emit_csr_kernel(const LabelEmitParams p) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const uint32_t tot = *p.n_entries;
    p.out[FR_OUT_LABEL_ENTRIES] = (float)tot;
    if (tot > p.cap) p.out[FR_OUT_OVERFLOW] = 1.f;
  }
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < p.S; r += gridDim.x * blockDim.x) {
    const int u = p.users[r / p.group];
    const int b = p.lab_off[u], cnt = p.lab_off[u + 1] - b;
    const float ws = p.ws_row[r];
    const uint32_t off = p.offs[r];
    for (int q = 0; q < cnt; ++q) {
      const uint32_t dst = off + q;
      if (dst < p.cap) { p.ent_key[dst] = (uint32_t)p.lab_idx[b + q]; p.ent_row[dst] = (uint32_t)r; p.ent_coef[dst] = ws; }
    }
  }
}

static int warp_grid(int nwarps_needed, int sm_count) {
  int grid = (nwarps_needed + FR_WARPS_PER_BLOCK - 1) / FR_WARPS_PER_BLOCK;
  if (grid > sm_count * 8) grid = sm_count * 8;
  return grid < 1 ? 1 : grid;
}
void launch_label_count(const LabelEmitParams& p, const Launch& l) {
  if (!p.user_labels) label_count_csr_kernel<<<max(1, min((p.S + 255) / 256, l.sm_count * 8)), 256, 0, l.st>>>(p);
  else label_count_kernel<<<warp_grid(p.S, l.sm_count), FR_THREADS, 0, l.st>>>(p);
  ++g_launches;
}
void launch_label_emit(const LabelEmitParams& p, const Launch& l) {
  if (!p.user_labels) label_emit_csr_kernel<<<max(1, min((p.S + 255) / 256, l.sm_count * 8)), 256, 0, l.st>>>(p);
  else label_emit_kernel<<<warp_grid(p.S, l.sm_count), FR_THREADS, 0, l.st>>>(p);
  ++g_launches;
}

// ------------------------------------------------------------------ Adam sweep / fill / mean
// Rows with last < target get the decay-only steps last+1..target (TF-1.x sparse Adam
// touches every row every step).  One thread per float4; stamps are rewritten afterwards
// by fill_i32 (separate launch: no intra-row race on `last`).
template <int OPT>
__global__ void __launch_bounds__(256)
adam_sweep_kernel(float4* __restrict__ var, float4* __restrict__ m, float4* __restrict__ v,
                  const int32_t* __restrict__ last, int64_t n4, int rowDV, const OptConsts oc, int target) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int lasti = last[i / rowDV];
    if (lasti >= target) continue;
    float4 mm = __ldcs(m + i), vv = __ldcs(v + i);
    if (mm.x == 0.f && mm.y == 0.f && mm.z == 0.f && mm.w == 0.f &&
        vv.x == 0.f && vv.y == 0.f && vv.z == 0.f && vv.w == 0.f) continue;
    float4 x = __ldcs(var + i);
    adam_catchup<OPT, 1>(&x, &mm, &vv, lasti, target, oc);     // DENSE and EXACT: step-by-step replay
    __stcs(var + i, x); __stcs(m + i, mm); __stcs(v + i, vv);
  }
}
// LAZY_SERIES coefficient table.  Invariant after step t: for every t0 < t
//   cser[5*t0 + n] = sum_{s=t0+1..t} lr_s * b1^(s-t0) * (1 - b2^((s-t0)/2))^n      (double)
// so a row stamped `last = t0` is brought to step t with one table row.  Each step adds the
// s = t term to the SERIES_WINDOW most recent rows (older rows' terms are < 1e-90).
__global__ void series_update_kernel(double* __restrict__ cser, const float* __restrict__ lr_hist, int t,
                                     double ln_b1, double half_ln_b2) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int t0 = t - j;
  if (j > SERIES_WINDOW || t0 < 0) return;
  double p = (double)lr_hist[t] * exp(j * ln_b1);
  const double d = -expm1(j * half_ln_b2);
  double* dst = cser + (size_t)t0 * SERIES_TERMS;
#pragma unroll
  for (int n = 0; n < SERIES_TERMS; ++n) { dst[n] += p; p *= d; }
}
__global__ void series_rebuild_kernel(double* __restrict__ cser, const float* __restrict__ lr_hist, int step,
                                      double ln_b1, double half_ln_b2) {
  const int t0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (t0 >= step) return;
  double acc[SERIES_TERMS];
#pragma unroll
  for (int n = 0; n < SERIES_TERMS; ++n) acc[n] = 0.0;
  const int jmax = min(step - t0, SERIES_WINDOW);
  for (int j = 1; j <= jmax; ++j) {
    double p = (double)lr_hist[t0 + j] * exp(j * ln_b1);
    const double d = -expm1(j * half_ln_b2);
#pragma unroll
    for (int n = 0; n < SERIES_TERMS; ++n) { acc[n] += p; p *= d; }
  }
#pragma unroll
  for (int n = 0; n < SERIES_TERMS; ++n) cser[(size_t)t0 * SERIES_TERMS + n] = acc[n];
}
void launch_series_update(double* cser, const float* lr_hist, int t, float b1, float b2, const Launch& l) {
  if (t < 1) return;
  series_update_kernel<<<SERIES_WINDOW / 256, 256, 0, l.st>>>(cser, lr_hist, t, log((double)b1), 0.5 * log((double)b2));
  ++g_launches;
}
void launch_series_rebuild(double* cser, const float* lr_hist, int step, float b1, float b2, const Launch& l) {
  if (step < 1) return;
  series_rebuild_kernel<<<(step + 255) / 256, 256, 0, l.st>>>(cser, lr_hist, step, log((double)b1), 0.5 * log((double)b2));
  ++g_launches;
}

__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t val) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = val;
}
void launch_adam_sweep(float4* var, float4* m, float4* v, int32_t* last, int64_t nrows, int rowDV,
                       const OptConsts& oc, int target_step, const Launch& l) {
  const int64_t n4 = nrows * rowDV;
  if (n4 == 0) return;
  int64_t grid = (n4 + 255) / 256;
  if (grid > (int64_t)l.sm_count * 16) grid = (int64_t)l.sm_count * 16;
  if (oc.adam_mode == FR_ADAM_LAZY_SERIES)
    adam_sweep_kernel<OPT_ADAM_SERIES><<<(int)grid, 256, 0, l.st>>>(var, m, v, last, n4, rowDV, oc, target_step);
  else
    adam_sweep_kernel<OPT_ADAM_EXACT><<<(int)grid, 256, 0, l.st>>>(var, m, v, last, n4, rowDV, oc, target_step);
  ++g_launches;
  launch_fill_i32(last, nrows, target_step, l);
}
void launch_fill_i32(int32_t* p, int64_t n, int32_t v, const Launch& l) {
  if (n == 0) return;
  int64_t grid = (n + 255) / 256;
  if (grid > (int64_t)l.sm_count * 8) grid = (int64_t)l.sm_count * 8;
  fill_i32_kernel<<<(int)grid, 256, 0, l.st>>>(p, n, v);
  ++g_launches;
}

// mean of a table (reduce_mean, :218-219): fixed grid, double partials, fixed-order final sum.
constexpr int MEAN_BLOCKS = 1024;
__global__ void __launch_bounds__(256)
mean_partial_kernel(const float4* __restrict__ x, int64_t n4, double* __restrict__ partials) {
  __shared__ double sh[8];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcs(x + i);
    s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
  }
  s = warp_sum_d(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    partials[blockIdx.x] = t;
  }
}
__global__ void mean_final_kernel(const double* partials, int nb, double count, float* out_slot) {
  double s = 0.0;
  for (int b = threadIdx.x; b < nb; b += 32) s += partials[b];
  s = warp_sum_d(s);
  if (threadIdx.x == 0) *out_slot = (float)(s / count);
}
void launch_mean(const float4* x, int64_t n4, double* partials, float* out_slot, double count, const Launch& l) {
  int64_t nb = (n4 + 255) / 256;
  if (nb > MEAN_BLOCKS) nb = MEAN_BLOCKS;
  if (nb < 1) nb = 1;
  mean_partial_kernel<<<(int)nb, 256, 0, l.st>>>(x, n4, partials);
  mean_final_kernel<<<1, 32, 0, l.st>>>(partials, (int)nb, count, out_slot);
  g_launches += 2;
}

__global__ void write_counters_kernel(const uint32_t* counters, float* out) {
  out[FR_OUT_UNIQ_USERS] = (float)counters[0];
  out[FR_OUT_UNIQ_ITEMS] = (float)counters[1];
}
void launch_write_counters(const uint32_t* counters, float* out, const Launch& l) {
  write_counters_kernel<<<1, 1, 0, l.st>>>(counters, out);
  ++g_launches;
}

}  // namespace fr
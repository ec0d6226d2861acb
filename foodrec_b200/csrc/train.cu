// Train-step kernels of the Recommender hot path (sm_100a).
//
//   fwd_train_kernel   Model_Recommender.py:56-104 + per-slice gradient norms (:236-237)
//   finalize_kernel    clip_by_global_norm scale (:237) + dense optimizer on Category_Embedding
//   seg_chunk_kernel / seg_combine_kernel
//                      sort-and-segment reduce of the sparse gradients fused with the
//                      optimizer (:240) and with Write_Memory (:106-220)
//   adam_sweep_kernel  TF-1.x non-lazy Adam decay of untouched rows (dense sweep / lazy flush)
//
// All kernels are HBM-bound row movers: one warp owns one table row at a time, a row
// of D floats is moved as 16-byte vectors (lane l <-> float4 l, l+32, ...), the four
// category rows live in shared memory.
#include "common.cuh"
#include "train.cuh"

namespace fr {

// ------------------------------------------------------------------ optimizer math
// Explicit _rn intrinsics: no FMA contraction, so the per-element arithmetic is the
// IEEE sequence the TF CPU kernels (and the numpy oracle) perform.
// Decay-only Adam steps (rows NOT in the batch, adam.py _apply_sparse_shared):
//   m <- b1*m ; v <- b2*v ; var <- var - lr_s*m/(sqrt(v)+eps)      for s = from..to
// One routine serves the dense sweep, the lazy catch-up and the lazy forward, so the
// three produce bit-identical rows.  All N float4 of the thread advance together inside
// one loop over steps (4N independent chains: throughput- not latency-bound); the
// quotient uses MUFU sqrt/rcp (<= 2 ulp on a term that is itself <= lr_s: far inside the
// 1e-5 parity bound) -- the loop is the only part of the step whose cost scales with the
// number of skipped steps.
__device__ __forceinline__ float fast_sqrt(float x) {
  // .ftz: one MUFU, no denormal fix-up code.  A denormal v flushes to sqrt = 0, and
  // 0 + eps == sqrt(v) + eps in fp32 for any v < 2^-126 (sqrt(v) < 1e-19 << ulp(eps)).
  float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}
__device__ __forceinline__ float fast_rcp(float x) {    // argument >= eps: always normal
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}
__device__ __forceinline__ void adam_decay1(float& var, float& m, float& v, float lr, const OptConsts& oc) {
  m = __fmul_rn(m, oc.b1);
  v = __fmul_rn(v, oc.b2);
  const float r = fast_rcp(__fadd_rn(fast_sqrt(v), oc.eps));
  var = __fmaf_rn(-__fmul_rn(lr, m), r, var);
}
template <int N>
__device__ __forceinline__ void adam_replay(float4* var, float4* m, float4* v, int from, int to,
                                            const OptConsts& oc) {
  if (from > to) return;
  bool any = false;
#pragma unroll
  for (int i = 0; i < N; ++i)
    any |= (m[i].x != 0.f) | (m[i].y != 0.f) | (m[i].z != 0.f) | (m[i].w != 0.f) |
           (v[i].x != 0.f) | (v[i].y != 0.f) | (v[i].z != 0.f) | (v[i].w != 0.f);
  if (!any) return;                      // never-touched elements: every skipped step is a no-op
  for (int s = from; s <= to; ++s) {
    const float lr = __ldg(oc.lr_hist + s);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      adam_decay1(var[i].x, m[i].x, v[i].x, lr, oc); adam_decay1(var[i].y, m[i].y, v[i].y, lr, oc);
      adam_decay1(var[i].z, m[i].z, v[i].z, lr, oc); adam_decay1(var[i].w, m[i].w, v[i].w, lr, oc);
    }
  }
}
__device__ __forceinline__ void adam_touch(float& var, float& m, float& v, float g, const OptConsts& oc) {
  m = __fadd_rn(__fmul_rn(m, oc.b1), __fmul_rn(g, oc.omb1));
  v = __fadd_rn(__fmul_rn(v, oc.b2), __fmul_rn(__fmul_rn(g, g), oc.omb2));
  var = __fsub_rn(var, __fdiv_rn(__fmul_rn(oc.lr_t, m), __fadd_rn(__fsqrt_rn(v), oc.eps)));
}
__device__ __forceinline__ void adagrad_touch(float& var, float& acc, float g, const OptConsts& oc) {
  acc = __fadd_rn(acc, __fmul_rn(g, g));
  var = __fsub_rn(var, __fdiv_rn(__fmul_rn(oc.lr, g), __fsqrt_rn(acc)));
}
__device__ __forceinline__ void rmsprop_touch(float& var, float& ms, float& mom, float g, const OptConsts& oc) {
  ms = __fadd_rn(ms, __fmul_rn(__fsub_rn(__fmul_rn(g, g), ms), oc.omrho));
  mom = __fadd_rn(__fmul_rn(mom, 0.f), __fdiv_rn(__fmul_rn(oc.lr, g), __fsqrt_rn(__fadd_rn(ms, oc.rms_eps))));
  var = __fsub_rn(var, mom);
}
__device__ __forceinline__ void sgd_touch(float& var, float g, const OptConsts& oc) {
  var = __fsub_rn(var, __fmul_rn(oc.lr, g));
}

#define FR_EACH4(V, ...) { { auto& X = V; (void)X; } __VA_ARGS__ }

// State of NR consecutive table rows (one "unique row" of P is 5 of them).
template <int NR, int NV>
struct RowState {
  float4 var[NR][NV], s1[NR][NV], s2[NR][NV];
  int last;
};

template <int NR, int NV>
__device__ __forceinline__ void load_state(RowState<NR, NV>& st, const float4* var_t, const float4* s1_t,
                                           const float4* s2_t, const int32_t* last_t, uint32_t rowid,
                                           const OptConsts& oc, int DV, int lane) {
  const size_t base = (size_t)rowid * NR * DV;
#pragma unroll
  for (int s = 0; s < NR; ++s)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      const bool ok = i < DV;
      const size_t off = base + (size_t)s * DV + i;
      st.var[s][k] = ok ? __ldcs(var_t + off) : f4zero();
      st.s1[s][k] = (ok && oc.learner != FR_SGD) ? __ldcs(s1_t + off) : f4zero();
      st.s2[s][k] = (ok && (oc.learner == FR_ADAM || oc.learner == FR_RMSPROP)) ? __ldcs(s2_t + off) : f4zero();
    }
  st.last = (oc.learner == FR_ADAM) ? last_t[rowid] : 0;
}

template <int NR, int NV>
__device__ __forceinline__ void apply_and_store(RowState<NR, NV>& st, float4* var_t, float4* s1_t, float4* s2_t,
                                                int32_t* last_t, uint32_t rowid, const float4 (&grad)[NR][NV],
                                                const OptConsts& oc, int DV, int lane,
                                                const float4 (*w1)[NV], const float4 (*w2)[NV], float alpha) {
  const size_t base = (size_t)rowid * NR * DV;
  if (oc.learner == FR_ADAM)      // lazy-exact: the steps this row sat out, then this step's update
    adam_replay<NR * NV>(&st.var[0][0], &st.s1[0][0], &st.s2[0][0], st.last + 1, oc.step - 1, oc);
#pragma unroll
  for (int s = 0; s < NR; ++s)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      if (i >= DV) continue;
      float4 var = st.var[s][k], a = st.s1[s][k], b = st.s2[s][k];
      const float4 g = grad[s][k];
      if (oc.learner == FR_ADAM) {
        adam_touch(var.x, a.x, b.x, g.x, oc); adam_touch(var.y, a.y, b.y, g.y, oc);
        adam_touch(var.z, a.z, b.z, g.z, oc); adam_touch(var.w, a.w, b.w, g.w, oc);
      } else if (oc.learner == FR_ADAGRAD) {
        adagrad_touch(var.x, a.x, g.x, oc); adagrad_touch(var.y, a.y, g.y, oc);
        adagrad_touch(var.z, a.z, g.z, oc); adagrad_touch(var.w, a.w, g.w, oc);
      } else if (oc.learner == FR_RMSPROP) {
        rmsprop_touch(var.x, a.x, b.x, g.x, oc); rmsprop_touch(var.y, a.y, b.y, g.y, oc);
        rmsprop_touch(var.z, a.z, b.z, g.z, oc); rmsprop_touch(var.w, a.w, b.w, g.w, oc);
      } else {
        sgd_touch(var.x, g.x, oc); sgd_touch(var.y, g.y, oc); sgd_touch(var.z, g.z, oc); sgd_touch(var.w, g.w, oc);
      }
      if (w1) {   // Write_Memory on personal steps: P += bias (:167), then P += alpha*general_bias (:198)
        var = add4(var, w1[s][k]);
        const float4 gb = scale4(alpha, w2[s][k]);
        var = add4(var, gb);
      }
      const size_t off = base + (size_t)s * DV + i;
      __stcs(var_t + off, var);
      if (oc.learner != FR_SGD) __stcs(s1_t + off, a);
      if (oc.learner == FR_ADAM || oc.learner == FR_RMSPROP) __stcs(s2_t + off, b);
    }
  if (oc.learner == FR_ADAM && lane == 0) last_t[rowid] = oc.step;
}

// ------------------------------------------------------------------ prep
__global__ void prep_rows_kernel(int mode, int B, const int32_t* __restrict__ users,
                                 const float* __restrict__ labels, const float* __restrict__ ws_in,
                                 uint32_t* __restrict__ ukeys, float* __restrict__ ws_row) {
  const int group = mode == FR_BPR ? 2 : 1;
  const int S = B * group;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < S; r += gridDim.x * blockDim.x) {
    const int grp = r / group;
    ukeys[r] = (uint32_t)users[grp];
    float w;
    if (ws_in) w = ws_in[r];
    else if (mode == FR_BPR) w = (r & 1) ? -1.f : 1.f;
    else w = labels[grp] > 0.5f ? 1.f : -1.f;
    ws_row[r] = w;
  }
}

void launch_prep_rows(int mode, int B, const int32_t* users, const float* labels, const float* ws_in,
                      uint32_t* ukeys, float* ws_row, const Launch& l) {
  const int S = B * (mode == FR_BPR ? 2 : 1);
  int grid = (S + 255) / 256;
  if (grid > l.sm_count * 8) grid = l.sm_count * 8;
  if (grid < 1) grid = 1;
  prep_rows_kernel<<<grid, 256, 0, l.st>>>(mode, B, users, labels, ws_in, ukeys, ws_row);
  ++g_launches;
}

// ------------------------------------------------------------------ forward + loss + norms
// One warp per group (sample / BPR triple).  Reads 5D floats of P[u] and D floats of
// R[i] per item row (24*D+32 bytes per sample), writes the z-row stash
// (1-a)/n * sum_c m_c P[u,1+c] so the recipe-gradient pass never re-reads P.
template <int NV, int GROUP>
__global__ void __launch_bounds__(FR_THREADS)
fwd_train_kernel(const FwdParams p) {
  extern __shared__ float4 smem[];
  const int DV = p.DV;
  float4* sCat = smem;                 // [4*DV]
  float4* red = smem + 4 * DV;         // [WARPS][4*DV]
  __shared__ float red_loss[FR_WARPS_PER_BLOCK], red_nrm[FR_WARPS_PER_BLOCK];
  for (int i = threadIdx.x; i < 4 * DV; i += blockDim.x) sCat[i] = p.cat[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * FR_WARPS_PER_BLOCK + warp, nw = gridDim.x * FR_WARPS_PER_BLOCK;
  const float a = p.a, oma = p.oma, Bf = (float)p.B;
  float4 gc[4][NV];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < NV; ++k) gc[c][k] = f4zero();
  float lossacc = 0.f, nrmacc = 0.f;

  for (int grp = gw; grp < p.B; grp += nw) {
    const int u = p.users[grp];
    if (grp + nw < p.B) {       // L2-prefetch the rows of this warp's next group
      const int gn = grp + nw;
      const uint32_t ub = 5u * (uint32_t)DV * 16u, rb = (uint32_t)DV * 16u;
      if (lane == 0) prefetch_l2_span(p.P + (size_t)p.users[gn] * 5 * DV, ub);
      else if (lane == 1 && p.lazy) prefetch_l2_span(p.mP + (size_t)p.users[gn] * 5 * DV, ub);
      else if (lane == 2 && p.lazy) prefetch_l2_span(p.vP + (size_t)p.users[gn] * 5 * DV, ub);
      else if (lane >= 3 && lane < 3 + GROUP) prefetch_l2_span(p.R + (size_t)p.items[gn * GROUP + lane - 3] * DV, rb);
    }
    float4 pr[5][NV];
#pragma unroll
    for (int s = 0; s < 5; ++s) load_row<NV>(pr[s], p.P + ((size_t)u * 5 + s) * DV, DV, lane);
    if (p.lazy) {
      const int from = p.lastP[u] + 1, to = p.oc.step - 1;
      if (from <= to) {
        float4 mm[5][NV], vv[5][NV];
#pragma unroll
        for (int s = 0; s < 5; ++s) {
          load_row<NV>(mm[s], p.mP + ((size_t)u * 5 + s) * DV, DV, lane);
          load_row<NV>(vv[s], p.vP + ((size_t)u * 5 + s) * DV, DV, lane);
        }
        adam_replay<5 * NV>(&pr[0][0], &mm[0][0], &vv[0][0], from, to, p.oc);
      }
    }
    float4 rr[GROUP][NV], pcn[GROUP][NV];   // R rows, pooledCat (normalised)
    float4 mm[GROUP];
    float sc[GROUP], nn[GROUP], nzq[GROUP], nRq[GROUP], npcq[GROUP];
#pragma unroll
    for (int j = 0; j < GROUP; ++j) {
      const int r = grp * GROUP + j;
      const int it = p.items[r];
      const float4 m = __ldg(p.cats + (p.cats_by_item ? it : r));
      const float n = ((m.x + m.y) + m.z) + m.w;                    // :77
      load_row_ro<NV>(rr[j], p.R + (size_t)it * DV, DV, lane);
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
        nR += dot4(rr[j][k], rr[j][k]);
        npc += dot4(pcs[k], pcs[k]);
      }
      hs = warp_sum(hs); ls = warp_sum(ls); nz = warp_sum(nz); nR = warp_sum(nR); npc = warp_sum(npc);
      const float high = hs / n, low = ls / n;                      // :79, :92
      sc[j] = a * high + oma * low;                                 // :95-96
      const float zc = oma / n;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = lane + 32 * k;
        if (i < DV) __stcg(p.z + (size_t)r * DV + i, scale4(zc, zs[k]));
        pcn[j][k] = div4(pcs[k], n);
      }
      mm[j] = m; nn[j] = n;
      const float inv2 = 1.f / (n * n);
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
      const float ga = g * a, n = nn[0];
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        fma4(gc[0][k], ga * (m.x / n), pr[0][k]); fma4(gc[1][k], ga * (m.y / n), pr[0][k]);
        fma4(gc[2][k], ga * (m.z / n), pr[0][k]); fma4(gc[3][k], ga * (m.w / n), pr[0][k]);
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
      const float n0 = nn[0], n1 = nn[GROUP - 1];
      const float4 w0 = make_float4(m0.x / n0, m0.y / n0, m0.z / n0, m0.w / n0);
      const float4 w1 = make_float4(m1.x / n1, m1.y / n1, m1.z / n1, m1.w / n1);
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
      dq = warp_sum(dq);
      const float nrm = h * h * (dq + oma * oma * (nzq[0] + nzq[GROUP - 1]));
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
  }
  // deterministic block reduction (fixed warp order), one partial per block
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      if (i < DV) red[(warp * 4 + c) * DV + i] = gc[c][k];
    }
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
  const int cap = sm_count * 3;      // 3 CTAs of 8 warps are resident per SM (<= 85 regs/thread)
  if (grid > cap) grid = cap;
  return grid < 1 ? 1 : grid;
}

void launch_fwd_train(int NV, int group, const FwdParams& p, int grid, const Launch& l) {
  const size_t smem = (size_t)(4 * p.DV) * sizeof(float4) * (1 + FR_WARPS_PER_BLOCK);
#define FR_FWD(NVV, GG) fwd_train_kernel<NVV, GG><<<grid, FR_THREADS, smem, l.st>>>(p)
  if (NV == 1) { if (group == 1) FR_FWD(1, 1); else FR_FWD(1, 2); }
  else         { if (group == 1) FR_FWD(2, 1); else FR_FWD(2, 2); }
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
      float acc = s1[j]; adagrad_touch(x, acc, g, oc); s1[j] = acc;
    } else if (oc.learner == FR_RMSPROP) {
      float ms = s1[j], mom = s2[j]; rmsprop_touch(x, ms, mom, g, oc); s1[j] = ms; s2[j] = mom;
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

// ------------------------------------------------------------------ segment reduce skeleton
// Entries are sorted by key (stable).  A warp owns CHUNK=32 consecutive entries and
// walks the runs of equal keys inside them in order.  A run fully inside the chunk is
// reduced and applied immediately (state rows prefetched before the reduction).  A run
// that crosses a chunk boundary leaves a partial ("piece") in slot 0 (run contains the
// chunk's first entry) or slot 1 (otherwise); seg_combine_kernel then sums a crossing
// run's pieces in chunk order -- deterministic, no atomics -- and applies it once.
template <class Pol>
__global__ void __launch_bounds__(FR_THREADS)
seg_chunk_kernel(const SegCommon c, const Pol pol) {
  extern __shared__ float4 smem[];
  constexpr int NR = Pol::NR, NV = Pol::NV;
  const int DV = pol.DV();
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
    {   // state rows of the second run of the chunk (the first is loaded right away)
      const uint32_t r1 = hm & ~1u;
      if (r1) pol.prefetch_state(__shfl_sync(FR_FULL, key, __ffs(r1) - 1), lane);
    }
    while (e0 < cnt) {
      const uint32_t rest = (e0 >= 31) ? 0u : (hm & ~((2u << e0) - 1u));
      const int e1 = rest ? (__ffs(rest) - 1) : cnt;
      {   // ... and, while this run is processed, of the run after the next one
        const uint32_t r2 = rest & (rest - 1u);
        if (r2) pol.prefetch_state(__shfl_sync(FR_FULL, key, __ffs(r2) - 1), lane);
      }
      const bool starts = (e0 > 0) || !from_prev;
      const bool ends = (e1 < cnt) || !to_next;
      const uint32_t k = __shfl_sync(FR_FULL, key, e0);
      float4 acc[NR][NV];
#pragma unroll
      for (int s = 0; s < NR; ++s)
#pragma unroll
        for (int q = 0; q < NV; ++q) acc[s][q] = f4zero();
      if (starts && ends) {
        typename Pol::State st;
        pol.load_state(st, k, lane);
        pol.accumulate(acc, e, e0, e1, lane, smem);
        pol.apply(st, k, acc, lane);
      } else {
        pol.accumulate(acc, e, e0, e1, lane, smem);
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
    // end of the run: first position > base+31 whose key differs (binary search: the
    // keys are sorted), so the piece loads below are independent and can be pipelined
    uint32_t lo = base + 32, hi = n;
    while (lo < hi) {
      const uint32_t mid = lo + ((hi - lo) >> 1);
      if (c.keys[mid] == lastKey) lo = mid + 1; else hi = mid;
    }
    const uint32_t kend = (lo - 1) >> 5;          // last chunk holding an entry of the run
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

// ---- policy: Personal_Memory rows.  grad slice of item row r (App. A.3):
//   dP[u,0]   += g*a*pooledCat_r ;  dP[u,1+c] += g*(1-a)*w_rc*R[i_r]
// PERSONAL additionally carries the Write_Memory terms (acc rows 5..9 = bias,
// 10..14 = sum of label-mean general memory).
template <int NVV, bool PERSONAL>
struct UserPol {
  static constexpr int NV = NVV;
  static constexpr int NR = PERSONAL ? 15 : 5;
  UserPolParams p;
  struct Entry { int item; float g; float4 m; float ws; int grp; };
  using State = RowState<5, NVV>;
  __device__ __forceinline__ int DV() const { return p.mc.DV; }
  __device__ __forceinline__ const float4* cat_src() const { return p.cat; }
  __device__ __forceinline__ Entry load_entry(uint32_t row, bool valid) const {
    Entry e; e.item = 0; e.g = 0.f; e.m = make_float4(1.f, 0.f, 0.f, 0.f); e.ws = 0.f; e.grp = 0;
    if (valid) {
      e.item = p.items[row];
      prefetch_l2_span(p.R + (size_t)e.item * p.mc.DV, (uint32_t)p.mc.DV * 16u);
      e.g = p.g[row] * p.out[FR_OUT_SCALE];
      e.m = __ldg(p.cats + (p.cats_by_item ? e.item : (int)row));
      e.ws = p.ws_row[row];
      e.grp = (int)row / p.group;
    }
    return e;
  }
  __device__ __forceinline__ void accumulate(float4 (&acc)[NR][NV], const Entry& e, int e0, int e1, int lane,
                                             const float4* sCat) const {
    const int DVv = p.mc.DV;
    // two recipe rows in flight per iteration (a BPR triple is exactly one such pair)
    for (int j0 = e0; j0 < e1; j0 += 2) {
     const int jn = (j0 + 1 < e1) ? j0 + 1 : j0;
     float4 rr2[2][NV];
     load_row_ro<NV>(rr2[0], p.R + (size_t)__shfl_sync(FR_FULL, e.item, j0) * DVv, DVv, lane);
     load_row_ro<NV>(rr2[1], p.R + (size_t)__shfl_sync(FR_FULL, e.item, jn) * DVv, DVv, lane);
#pragma unroll
     for (int u = 0; u < 2; ++u) {
      const int j = j0 + u;
      if (j >= e1) break;
      const float g = __shfl_sync(FR_FULL, e.g, j);
      const float4 m = shfl4(e.m, j);
      float4 (&rr)[NV] = rr2[u];
      float4 pcs[NV];
      pooled_cat<NV>(pcs, sCat, m, DVv, lane);
      const float n = ((m.x + m.y) + m.z) + m.w;
      const float ga = g * p.mc.a, go = g * p.mc.oma;
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const float4 pc = div4(pcs[k], n);
        mad4_rn(acc[0][k], ga, pc);
        mad4_rn(acc[1][k], go * (m.x / n), rr[k]); mad4_rn(acc[2][k], go * (m.y / n), rr[k]);
        mad4_rn(acc[3][k], go * (m.z / n), rr[k]); mad4_rn(acc[4][k], go * (m.w / n), rr[k]);
        if (PERSONAL) {
          const float ws = __shfl_sync(FR_FULL, e.ws, j);
          const float hc = p.mc.beta_2 * ws, lc = p.mc.beta_1 * ws;
          mad4_rn(acc[5][k], hc, pc);                                   // :140-144
          mad4_rn(acc[6][k], lc, scale4(m.x, rr[k])); mad4_rn(acc[7][k], lc, scale4(m.y, rr[k]));   // :111-119
          mad4_rn(acc[8][k], lc, scale4(m.z, rr[k])); mad4_rn(acc[9][k], lc, scale4(m.w, rr[k]));
        }
      }
      if (PERSONAL) {            // (sum_l lam_l G_old[l]) / sum_l lam_l   (:170-186)
        const int grp = __shfl_sync(FR_FULL, e.grp, j);
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
          for (int k = 0; k < NV; ++k) acc[10 + s][k] = add4(acc[10 + s][k], div4(gs[s][k], lsum));
      }
     }
    }
  }
  __device__ __forceinline__ void load_state(State& st, uint32_t key, int lane) const {
    fr::load_state<5, NV>(st, p.P, p.s1, p.s2, p.last, key, p.oc, p.mc.DV, lane);
  }
  __device__ __forceinline__ void prefetch_state(uint32_t key, int lane) const {
    const uint32_t bytes = 5u * (uint32_t)p.mc.DV * 16u;      // a user's 5 slots are contiguous
    const size_t off = (size_t)key * 5 * p.mc.DV;
    const int L = p.oc.learner;
    if (lane == 0) prefetch_l2_span(p.P + off, bytes);
    else if (lane == 1 && L != FR_SGD) prefetch_l2_span(p.s1 + off, bytes);
    else if (lane == 2 && (L == FR_ADAM || L == FR_RMSPROP)) prefetch_l2_span(p.s2 + off, bytes);
  }
  __device__ __forceinline__ void apply(State& st, uint32_t key, float4 (&acc)[NR][NV], int lane) const {
    float4 (&grad)[5][NV] = *reinterpret_cast<float4 (*)[5][NV]>(&acc[0]);
    if (PERSONAL)
      apply_and_store<5, NV>(st, p.P, p.s1, p.s2, p.last, key, grad, p.oc, p.mc.DV, lane,
                             &acc[5], &acc[10], p.mc.alpha);
    else
      apply_and_store<5, NV>(st, p.P, p.s1, p.s2, p.last, key, grad, p.oc, p.mc.DV, lane,
                             nullptr, nullptr, 0.f);
  }
};

// ---- policy: Recipe_Embedding rows.  dR[i] += g * z_r  (z stashed by the forward pass)
template <int NVV>
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
    if (valid) e.g = p.g[row] * p.out[FR_OUT_SCALE];
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
    fr::load_state<1, NV>(st, p.R, p.s1, p.s2, p.last, key, p.oc, p.mc.DV, lane);
  }
  __device__ __forceinline__ void prefetch_state(uint32_t key, int lane) const {
    const uint32_t bytes = (uint32_t)p.mc.DV * 16u;
    const size_t off = (size_t)key * p.mc.DV;
    const int L = p.oc.learner;
    if (lane == 0) prefetch_l2_span(p.R + off, bytes);
    else if (lane == 1 && L != FR_SGD) prefetch_l2_span(p.s1 + off, bytes);
    else if (lane == 2 && (L == FR_ADAM || L == FR_RMSPROP)) prefetch_l2_span(p.s2 + off, bytes);
  }
  __device__ __forceinline__ void apply(State& st, uint32_t key, float4 (&acc)[1][NV], int lane) const {
    apply_and_store<1, NV>(st, p.R, p.s1, p.s2, p.last, key, acc, p.oc, p.mc.DV, lane, nullptr, nullptr, 0.f);
  }
};

// ---- policy: General_Memory rows (Write_Memory :201-215), entries = non-zeros of the
// label feed sorted by label:  G[l] += sum lam*ws*[beta_2*pooledCat ; beta_1*m_c*R[i]]
template <int NVV>
struct LabelPol {
  static constexpr int NV = NVV;
  static constexpr int NR = 5;
  LabelPolParams p;
  struct Entry { int item; float coef; float4 m; };
  struct State { float4 var[5][NVV]; };
  __device__ __forceinline__ int DV() const { return p.mc.DV; }
  __device__ __forceinline__ const float4* cat_src() const { return p.cat; }
  __device__ __forceinline__ Entry load_entry(uint32_t ent, bool valid) const {
    Entry e; e.item = 0; e.coef = 0.f; e.m = make_float4(1.f, 0.f, 0.f, 0.f);
    if (valid) {
      const uint32_t row = p.ent_row[ent];
      e.item = p.items[row];
      prefetch_l2_span(p.R + (size_t)e.item * p.mc.DV, (uint32_t)p.mc.DV * 16u);
      e.coef = p.ent_coef[ent];
      e.m = __ldg(p.cats + (p.cats_by_item ? e.item : (int)row));
    }
    return e;
  }
  __device__ __forceinline__ void accumulate(float4 (&acc)[5][NV], const Entry& e, int e0, int e1, int lane,
                                             const float4* sCat) const {
    const int DVv = p.mc.DV;
    constexpr int PF = 4;                       // recipe rows in flight
    for (int j0 = e0; j0 < e1; j0 += PF) {
      float4 rr4[PF][NV];
#pragma unroll
      for (int u = 0; u < PF; ++u)
        load_row_ro<NV>(rr4[u], p.R + (size_t)__shfl_sync(FR_FULL, e.item, min(j0 + u, e1 - 1)) * DVv, DVv, lane);
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int j = j0 + u;
        if (j >= e1) break;
        const float coef = __shfl_sync(FR_FULL, e.coef, j);
        const float4 m = shfl4(e.m, j);
        float4 pcs[NV];
        pooled_cat<NV>(pcs, sCat, m, DVv, lane);
        const float n = ((m.x + m.y) + m.z) + m.w;
        const float hc = p.mc.beta_2 * coef, lc = p.mc.beta_1 * coef;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
          mad4_rn(acc[0][k], hc, div4(pcs[k], n));
          mad4_rn(acc[1][k], lc, scale4(m.x, rr4[u][k])); mad4_rn(acc[2][k], lc, scale4(m.y, rr4[u][k]));
          mad4_rn(acc[3][k], lc, scale4(m.z, rr4[u][k])); mad4_rn(acc[4][k], lc, scale4(m.w, rr4[u][k]));
        }
      }
    }
  }
  __device__ __forceinline__ void prefetch_state(uint32_t, int) const {}   // G is tiny and hot
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

template <class Pol>
static void launch_seg(const SegCommon& c, const Pol& pol, int DV, bool needs_cat, const Launch& l) {
  const uint32_t nchunks = (c.n_host + 31) / 32;
  if (nchunks == 0) return;
  int grid = (int)((nchunks + FR_WARPS_PER_BLOCK - 1) / FR_WARPS_PER_BLOCK);
  const size_t smem = needs_cat ? (size_t)4 * DV * sizeof(float4) : 0;
  // persistent grid = exactly the resident CTAs (SM count x occupancy): no partial wave
  static int occ_chunk = 0;
  if (!occ_chunk) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_chunk, seg_chunk_kernel<Pol>, FR_THREADS, smem);
    if (occ_chunk < 1) occ_chunk = 1;
  }
  const int cap = l.sm_count * occ_chunk;
  if (grid > cap) grid = cap;
  seg_chunk_kernel<Pol><<<grid, FR_THREADS, smem, l.st>>>(c, pol);
  if (l.mid) cudaEventRecord(l.mid, l.st);
  seg_combine_kernel<Pol><<<grid, FR_THREADS, 0, l.st>>>(c, pol);
  g_launches += 2;
}

void launch_user_pass(int NV, int personal, const SegCommon& c, const UserPolParams& p, const Launch& l) {
  if (NV == 1) {
    if (personal) { UserPol<1, true> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); }
    else { UserPol<1, false> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); }
  } else {
    if (personal) { UserPol<2, true> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); }
    else { UserPol<2, false> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); }
  }
}
void launch_item_pass(int NV, const SegCommon& c, const ItemPolParams& p, const Launch& l) {
  if (NV == 1) { ItemPol<1> pol{p}; launch_seg(c, pol, p.mc.DV, false, l); }
  else { ItemPol<2> pol{p}; launch_seg(c, pol, p.mc.DV, false, l); }
}
void launch_label_pass(int NV, const SegCommon& c, const LabelPolParams& p, const Launch& l) {
  if (NV == 1) { LabelPol<1> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); }
  else { LabelPol<2> pol{p}; launch_seg(c, pol, p.mc.DV, true, l); }
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
    p.out[FR_OUT_OVERFLOW] = tot > p.cap ? 1.f : 0.f;
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

static int warp_grid(int nwarps_needed, int sm_count) {
  int grid = (nwarps_needed + FR_WARPS_PER_BLOCK - 1) / FR_WARPS_PER_BLOCK;
  if (grid > sm_count * 8) grid = sm_count * 8;
  return grid < 1 ? 1 : grid;
}
void launch_label_count(const LabelEmitParams& p, const Launch& l) {
  label_count_kernel<<<warp_grid(p.S, l.sm_count), FR_THREADS, 0, l.st>>>(p);
  ++g_launches;
}
void launch_label_emit(const LabelEmitParams& p, const Launch& l) {
  label_emit_kernel<<<warp_grid(p.S, l.sm_count), FR_THREADS, 0, l.st>>>(p);
  ++g_launches;
}

// ------------------------------------------------------------------ Adam sweep / fill / mean
// Rows with last < target get the decay-only steps last+1..target (TF-1.x sparse Adam
// touches every row every step).  One thread per float4; stamps are rewritten afterwards
// by fill_i32 (separate launch: no intra-row race on `last`).
__global__ void __launch_bounds__(256)
adam_sweep_kernel(float4* __restrict__ var, float4* __restrict__ m, float4* __restrict__ v,
                  const int32_t* __restrict__ last, int64_t n4, int rowDV, const OptConsts oc, int target) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int from = last[i / rowDV] + 1;
    if (from > target) continue;
    float4 mm = __ldcs(m + i), vv = __ldcs(v + i);
    if (mm.x == 0.f && mm.y == 0.f && mm.z == 0.f && mm.w == 0.f &&
        vv.x == 0.f && vv.y == 0.f && vv.z == 0.f && vv.w == 0.f) continue;
    float4 x = __ldcs(var + i);
    adam_replay<1>(&x, &mm, &vv, from, target, oc);
    __stcs(var + i, x); __stcs(m + i, mm); __stcs(v + i, vv);
  }
}
// lazy-exact Adam: the batch's unique recipe rows are brought to step-1 before anything
// reads them (forward, dP accumulation, Write_Memory all read R).  keys = recipe ids of the
// item rows, sorted; the warp that sees a run's head owns that row.
template <int NV>
__global__ void __launch_bounds__(FR_THREADS)
item_catchup_kernel(const uint32_t* __restrict__ keys, uint32_t n, float4* __restrict__ R,
                    float4* __restrict__ m, float4* __restrict__ v, int32_t* __restrict__ last,
                    int DV, const OptConsts oc) {
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
      const int from = last[k] + 1;
      __syncwarp();
      if (from > to) continue;
      float4 x[NV], mm[NV], vv[NV];
      load_row<NV>(mm, m + (size_t)k * DV, DV, lane);
      load_row<NV>(vv, v + (size_t)k * DV, DV, lane);
      load_row<NV>(x, R + (size_t)k * DV, DV, lane);
      adam_replay<NV>(x, mm, vv, from, to, oc);
      store_row<NV>(R + (size_t)k * DV, x, DV, lane);
      store_row<NV>(m + (size_t)k * DV, mm, DV, lane);
      store_row<NV>(v + (size_t)k * DV, vv, DV, lane);
      if (lane == 0) last[k] = to;
    }
  }
}
void launch_item_catchup(int NV, const uint32_t* keys, uint32_t n, float4* R, float4* m, float4* v,
                         int32_t* last, int DV, const OptConsts& oc, const Launch& l) {
  const uint32_t nchunks = (n + 31) / 32;
  if (nchunks == 0) return;
  int grid = (int)((nchunks + FR_WARPS_PER_BLOCK - 1) / FR_WARPS_PER_BLOCK);
  if (grid > l.sm_count * 16) grid = l.sm_count * 16;
  if (NV == 1) item_catchup_kernel<1><<<grid, FR_THREADS, 0, l.st>>>(keys, n, R, m, v, last, DV, oc);
  else item_catchup_kernel<2><<<grid, FR_THREADS, 0, l.st>>>(keys, n, R, m, v, last, DV, oc);
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
  adam_sweep_kernel<<<(int)grid, 256, 0, l.st>>>(var, m, v, last, n4, rowDV, oc, target_step);
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

// Optimizer arithmetic shared by the train kernels (TF-1.x semantics).
//
// Kernels are specialised at COMPILE time by `OPT` so that each one carries the code of
// exactly one optimizer / catch-up scheme: with everything behind run-time switches the
// user-pass kernel was > 64 KB of SASS and its top stall reason was instruction fetch
// (ncu: stalled_no_instruction), not memory.
#pragma once
#include "common.cuh"
#include "internal.h"

namespace fr {

enum { OPT_GENERIC = 0,        // SGD | Adagrad | RMSProp, selected at run time (small bodies)
       OPT_ADAM_DENSE = 1,     // TF-1.x literal: every row is current, no per-row catch-up
       OPT_ADAM_EXACT = 2,     // lazy, skipped steps replayed one by one
       OPT_ADAM_SERIES = 3 };  // lazy, skipped steps in closed form

__host__ __device__ inline int opt_of(int learner, int adam_mode) {
  if (learner != FR_ADAM) return OPT_GENERIC;
  return adam_mode == FR_ADAM_DENSE ? OPT_ADAM_DENSE
       : (adam_mode == FR_ADAM_LAZY_EXACT ? OPT_ADAM_EXACT : OPT_ADAM_SERIES);
}

// ---- decay-only Adam steps (rows NOT in the batch; adam.py _apply_sparse_shared):
//   m <- b1*m ; v <- b2*v ; var <- var - lr_s*m/(sqrt(v)+eps)      for s = from..to
// One routine serves the dense sweep and the LAZY_EXACT catch-up, so both produce
// bit-identical rows.  All N float4 of the thread advance together inside one loop over
// steps (4N independent chains); the quotient uses MUFU sqrt/rcp (<= 2 ulp on a term that
// is itself <= lr_s: far inside the 1e-5 parity bound).
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

// ---- LAZY_SERIES: the same k decay-only steps in closed form (see foodrec_b200.h).  With
// q = sqrt(v), d = q + eps, y = q/d:   var -= (m/d) * sum_n C_n y^n ;  m *= b1^k ; v *= b2^k.
// The C_n depend only on (last, now) and are read once per row from the table.
struct SeriesRow { float c[SERIES_TERMS]; float p1, p2; };
__device__ __forceinline__ SeriesRow series_row(const OptConsts& oc, int last, int now) {
  SeriesRow r;
  const double* src = oc.cser + (size_t)last * SERIES_TERMS;
#pragma unroll
  for (int n = 0; n < SERIES_TERMS; ++n) r.c[n] = (float)__ldg(src + n);
  const float k = (float)(now - last);
  r.p1 = exp2f(k * oc.l2b1);
  r.p2 = exp2f(k * oc.l2b2);
  return r;
}
__device__ __forceinline__ void adam_series1(float& var, float& m, float& v, const SeriesRow& r, float eps) {
  const float q = fast_sqrt(v);
  float rd = fast_rcp(__fadd_rn(q, eps));
  rd = rd * fmaf(-__fadd_rn(q, eps), rd, 2.0f);          // one Newton step: full fp32 accuracy
  const float y = q * rd;
  float poly = r.c[SERIES_TERMS - 1];
#pragma unroll
  for (int n = SERIES_TERMS - 2; n >= 0; --n) poly = fmaf(poly, y, r.c[n]);
  var = fmaf(-(m * rd), poly, var);
  m *= r.p1;
  v *= r.p2;
}

// Bring N float4 of one row from step `last` to step `now` (decay-only steps last+1..now).
template <int OPT, int N>
__device__ __forceinline__ void adam_catchup(float4* var, float4* m, float4* v, int last, int now,
                                             const OptConsts& oc) {
  if (last >= now) return;
  if constexpr (OPT == OPT_ADAM_SERIES) {
    const SeriesRow r = series_row(oc, last, now);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      adam_series1(var[i].x, m[i].x, v[i].x, r, oc.eps); adam_series1(var[i].y, m[i].y, v[i].y, r, oc.eps);
      adam_series1(var[i].z, m[i].z, v[i].z, r, oc.eps); adam_series1(var[i].w, m[i].w, v[i].w, r, oc.eps);
    }
  } else {
    adam_replay<N>(var, m, v, last + 1, now, oc);
  }
}

// ---- the step's own update of a row that IS in the batch.  Explicit _rn intrinsics: no FMA
// contraction, i.e. the IEEE sequence of the TF CPU kernels and of the numpy oracle.
// m and v follow the TF op order exactly (they feed Adam's ill-conditioned quotient); the
// quotient itself uses MUFU sqrt + rcp with one Newton step (<= 2 ulp of a term <= lr_t)
// instead of IEEE div/sqrt: 20 of these are inlined per user row and code size matters.
__device__ __forceinline__ void adam_touch(float& var, float& m, float& v, float g, const OptConsts& oc) {
  m = __fadd_rn(__fmul_rn(m, oc.b1), __fmul_rn(g, oc.omb1));
  v = __fadd_rn(__fmul_rn(v, oc.b2), __fmul_rn(__fmul_rn(g, g), oc.omb2));
  const float d = __fadd_rn(fast_sqrt(v), oc.eps);
  float r = fast_rcp(d);
  r = r * fmaf(-d, r, 2.0f);
  var = fmaf(-__fmul_rn(oc.lr_t, m), r, var);
}
// TF-1.15 core/kernels/training_ops.cc uses DIFFERENT arithmetic forms for its sparse and its dense apply ops
// (oracle/recommender_oracle.py:_adagrad/_rmsprop restate both): SPARSE = the rows of P and R (SparseApply*Op),
// dense = Category_Embedding (Apply*<CPUDevice>, finalize_kernel).  rsqrt(x) is the IEEE 1/sqrt(x).
template <bool SPARSE>
__device__ __forceinline__ void adagrad_touch(float& var, float& acc, float g, const OptConsts& oc) {
  acc = __fadd_rn(acc, __fmul_rn(g, g));                       // a += g.square()
  const float rs = __frcp_rn(__fsqrt_rn(acc));
  // sparse: v -= (lr * g) * a.rsqrt();  dense: var -= (grad * lr) * accum.rsqrt()  (the products commute)
  var = __fsub_rn(var, __fmul_rn(__fmul_rn(oc.lr, g), rs));
}
template <bool SPARSE>
__device__ __forceinline__ void rmsprop_touch(float& var, float& ms, float& mom, float g, const OptConsts& oc) {
  if constexpr (SPARSE) {
    // ms = ms * rho + grad.square() * (1 - rho);  mom = mom * momentum + (ms + eps).rsqrt() * lr * grad
    ms = __fadd_rn(__fmul_rn(ms, oc.rho), __fmul_rn(__fmul_rn(g, g), oc.omrho));
    const float rs = __frcp_rn(__fsqrt_rn(__fadd_rn(ms, oc.rms_eps)));
    mom = __fadd_rn(__fmul_rn(mom, 0.f), __fmul_rn(__fmul_rn(rs, oc.lr), g));
  } else {
    // ms += (grad.square() - ms) * (1 - rho);  mom = mom * momentum + (grad * lr) / (ms + eps).sqrt()
    ms = __fadd_rn(ms, __fmul_rn(__fsub_rn(__fmul_rn(g, g), ms), oc.omrho));
    mom = __fadd_rn(__fmul_rn(mom, 0.f), __fdiv_rn(__fmul_rn(g, oc.lr), __fsqrt_rn(__fadd_rn(ms, oc.rms_eps))));
  }
  var = __fsub_rn(var, mom);
}
__device__ __forceinline__ void sgd_touch(float& var, float g, const OptConsts& oc) {
  var = __fsub_rn(var, __fmul_rn(oc.lr, g));
}

// State of NR consecutive table rows (one "unique row" of P is 5 of them).
template <int NR, int NV>
struct RowState {
  float4 var[NR][NV], s1[NR][NV], s2[NR][NV];
  int last;
};

// BF: the variable table is stored in bf16 (fr_set_table_format); the optimizer slots are always fp32.
template <int OPT, int NR, int NV, bool BF = false>
__device__ __forceinline__ void load_state(RowState<NR, NV>& st, const float4* var_t, const float4* s1_t,
                                           const float4* s2_t, const int32_t* last_t, uint32_t rowid,
                                           const OptConsts& oc, int DV, int lane) {
  const size_t base = (size_t)rowid * NR * DV;
  const bool has1 = OPT != OPT_GENERIC || oc.learner != FR_SGD;
  const bool has2 = OPT != OPT_GENERIC || oc.learner == FR_RMSPROP;
#pragma unroll
  for (int s = 0; s < NR; ++s)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      const bool ok = i < DV;
      const size_t off = base + (size_t)s * DV + i;
      st.var[s][k] = ok ? tab_ld_cs(tab_at<BF>(var_t, off)) : f4zero();
      st.s1[s][k] = (ok && has1) ? __ldcs(s1_t + off) : f4zero();
      st.s2[s][k] = (ok && has2) ? __ldcs(s2_t + off) : f4zero();
    }
  st.last = (OPT == OPT_ADAM_EXACT || OPT == OPT_ADAM_SERIES) ? last_t[rowid] : 0;
}

template <int OPT, int NR, int NV, bool BF = false>
__device__ __forceinline__ void apply_and_store(RowState<NR, NV>& st, float4* var_t, float4* s1_t, float4* s2_t,
                                                int32_t* last_t, uint32_t rowid, const float4 (&grad)[NR][NV],
                                                const OptConsts& oc, int DV, int lane) {
  const size_t base = (size_t)rowid * NR * DV;
  if constexpr (OPT == OPT_ADAM_EXACT || OPT == OPT_ADAM_SERIES)   // the steps this row sat out
    adam_catchup<OPT, NR * NV>(&st.var[0][0], &st.s1[0][0], &st.s2[0][0], st.last, oc.step - 1, oc);
  const bool has1 = OPT != OPT_GENERIC || oc.learner != FR_SGD;
  const bool has2 = OPT != OPT_GENERIC || oc.learner == FR_RMSPROP;
#pragma unroll
  for (int s = 0; s < NR; ++s)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      if (i >= DV) continue;
      float4 var = st.var[s][k], a = st.s1[s][k], b = st.s2[s][k];
      const float4 g = grad[s][k];
      if constexpr (OPT != OPT_GENERIC) {
        adam_touch(var.x, a.x, b.x, g.x, oc); adam_touch(var.y, a.y, b.y, g.y, oc);
        adam_touch(var.z, a.z, b.z, g.z, oc); adam_touch(var.w, a.w, b.w, g.w, oc);
      } else if (oc.learner == FR_ADAGRAD) {
        adagrad_touch<true>(var.x, a.x, g.x, oc); adagrad_touch<true>(var.y, a.y, g.y, oc);
        adagrad_touch<true>(var.z, a.z, g.z, oc); adagrad_touch<true>(var.w, a.w, g.w, oc);
      } else if (oc.learner == FR_RMSPROP) {
        rmsprop_touch<true>(var.x, a.x, b.x, g.x, oc); rmsprop_touch<true>(var.y, a.y, b.y, g.y, oc);
        rmsprop_touch<true>(var.z, a.z, b.z, g.z, oc); rmsprop_touch<true>(var.w, a.w, b.w, g.w, oc);
      } else {
        sgd_touch(var.x, g.x, oc); sgd_touch(var.y, g.y, oc); sgd_touch(var.z, g.z, oc); sgd_touch(var.w, g.w, oc);
      }
      const size_t off = base + (size_t)s * DV + i;
      tab_st_cs(tab_at<BF>(var_t, off), var);         // (bf16 table: round to nearest even on store)
      if (has1) __stcs(s1_t + off, a);
      if (has2) __stcs(s2_t + off, b);
    }
  if (OPT != OPT_GENERIC && lane == 0) last_t[rowid] = oc.step;
}

}  // namespace fr

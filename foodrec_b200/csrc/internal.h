// Internal declarations shared by the translation units of libfoodrec_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/foodrec_b200.h"

namespace fr {

constexpr int SORT_TILE = 2048;     // keys per block-iteration of the radix passes
constexpr int RADIX_BITS = 10;      // widest digit of a pass (sort.cu picks <= this per job)
constexpr int RADIX_BINS = 1 << RADIX_BITS;

struct SortBufs {
  uint32_t* k[2] = {nullptr, nullptr};
  uint32_t* v[2] = {nullptr, nullptr};
  uint32_t* tile_hist = nullptr;   // [RADIX_BINS * ntiles]
  uint32_t* scan_tmp = nullptr;    // block sums for the multi-block scan
  uint32_t* ticket = nullptr;      // [1], zero between launches: "last block done" counter of the one-launch offset scan
  int cap = 0;
};

// Stable LSD radix sort of (key, index).  keys_in is read by pass 0 (never written);
// result lands in bufs.k[r], bufs.v[r] with r returned.  n_dev (optional) overrides
// n_host with a device-resident count <= n_host.
int radix_sort_pairs(SortBufs& bufs, const uint32_t* keys_in, uint32_t n_host,
                     const uint32_t* n_dev, int nbits, cudaStream_t st, int sm_count);

// Up to two independent sorts sharing their launches (blockIdx.y): job i's output lands in
// (bufs->k[result], bufs->v[result]).
struct SortJob {
  SortBufs* bufs; const uint32_t* keys_in; uint32_t n_host; const uint32_t* n_dev; int nbits; int result;
};
void radix_sort_jobs(SortJob* jobs, int njobs, cudaStream_t st, int sm_count);

void exclusive_scan_u32(const uint32_t* in, uint32_t* out, uint32_t n, uint32_t* tmp,
                        uint32_t* total_out /*nullable*/, cudaStream_t st);

// ---- optimizer constants passed by value to the update kernels
struct OptConsts {
  int learner, adam_mode;
  float lr;            // args.lr
  float lr_t;          // Adam: lr*sqrt(1-b2^t)/(1-b1^t) of THIS step
  float b1, b2, eps;   // Adam
  float omb1, omb2;    // 1-b1, 1-b2 in fp32
  float rho, omrho, rms_eps;
  int step;            // t of this step (1-based)
  const float* lr_hist;  // lr_t of every past step, index = step
  // LAZY_SERIES: cser[5*t0 + n] = C_n(t0 -> current step count), see foodrec_b200.h
  const double* cser;
  float l2b1, l2b2;    // log2(beta1), log2(beta2)
};

// single-pass training (train_seg.cu, user_fused_kernel): bit 30 of a Personal_Memory stamp says which copy of the
// double-buffered row is current (0: the caller's table, 1: the shadow of fr_set_shadow)
constexpr uint32_t FR_SHADOW_BIT = 1u << 30;

constexpr int SERIES_TERMS = 5;
constexpr int SERIES_WINDOW = 2048;   // TF default betas: b1^j underflows fp32 long before j = 2048; fr_create checks the configured ones

struct ModelConsts {
  int D, DV, L;
  float a, oma;               // high coefficient, 1-a (fp32)
  float beta_1, beta_2, alpha;
};

// ---- per-step device views
struct StepView {
  // tables
  float4 *P, *R, *G;
  float4 *s1P, *s2P, *s1R, *s2R;
  int32_t *lastP, *lastR;
  const float4* cat_pre;   // snapshot of Cat taken at step start [4*DV]
  // batch (device)
  int mode, B, S, group;
  const int32_t *users, *items;
  const float4* cats; int cats_by_item;
  const float* labels;
  const float* write_sign;    // nullable -> derived
  const float* user_labels;   // dense [B,L] or null
  const int32_t *lab_off, *lab_idx;
  // per-row scratch
  float* g;        // [S] dL/ds per item row (unscaled by clip)
  float4* z;       // [S*DV] (1-a) * sum_c w_c P[u,1+c]
  float* scores;   // [S]
  float* out;      // device scalars FR_OUT_*
};

extern unsigned long long g_launches;   // kernels launched by this library (bench.py's gpu_launches)

struct Launch { int sm_count; cudaStream_t st; cudaEvent_t mid = nullptr; /* recorded between chunk and combine */
                const cudaAccessPolicyWindow* win = nullptr; /* persisting-L2 window for the launch (kernels that take one) */ };

// Health term at inference (fr_set_health_blend): the user row is scored as
//   P'[u] = P[u] + alpha * (sum_{l in labels(u)} G[l]) / |labels(u)|
// i.e. what Write_Memory materialises on a personal step (Model_Recommender.py:170-198), without writing it.
// G == nullptr: off.  Products and sums are rounded separately (no FMA) so the host restatement is bit-exact.
struct HealthBlend { const float4* G; const int32_t* lab_off; const int32_t* lab_idx; float alpha; };

void launch_fwd_score(const ModelConsts& mc, const float4* P, const float4* R, const float4* Cat,
                      const int32_t* users, const int32_t* items, const float4* cats,
                      int cats_by_item, int n, float* scores, const HealthBlend& hb, const Launch& l,
                      int64_t n_users, int64_t n_items, int bf16);

void launch_eval_sampled(const ModelConsts& mc, const float4* P, const float4* R, const float4* Cat,
                         const int32_t* users, const int32_t* cand, const int32_t* n_cand, int n_users,
                         int stride, const float4* cand_cats, const float4* item_cats, int K,
                         int32_t* topk_ids, int32_t* gt_rank, float* scores, const HealthBlend& hb, const Launch& l,
                         int64_t n_table_users, int64_t n_items, int bf16);

}  // namespace fr

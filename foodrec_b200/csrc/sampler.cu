// Negative sampler for 1:N BPR training (BASELINE configs[4]: 1:8 negatives).  The reference has no
// sampler (get_train_instances takes listed negatives, Train_recommender.py:86-93); this is the
// counter-based equivalent so that host and device draw the SAME negatives for the same (seed, sample
// index): Philox4x32-10 (Salmon et al., SC'11; the generator curand and torch use), one counter per
// (sample, negative, attempt), uniform over [0, I) by multiply-shift, positives rejected.
//   counter = (sample_lo, sample_hi, negative j, attempt), key = (seed_lo, seed_hi)
//   item    = mulhi32(x0, I);  item == positive -> next attempt (<= 16, then (positive + 1) % I)
// oracle/sampler_oracle.py restates it in numpy; ids are compared bit for bit.
#include "common.cuh"
#include "ctx.h"

namespace fr {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
    const uint32_t hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += W0; k.y += W1;
  }
  return c;
}

__global__ void __launch_bounds__(256)
sample_negatives_kernel(const int32_t* __restrict__ pos, long long n, int n_neg, uint32_t I, uint2 key,
                        unsigned long long offset, int32_t* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * n_neg) return;
  const long long s = t / n_neg;
  const uint32_t j = (uint32_t)(t % n_neg);
  const unsigned long long g = offset + (unsigned long long)s;
  const uint32_t p = (uint32_t)pos[s];
  uint32_t item = (p + 1u) % I;
  for (uint32_t attempt = 0; attempt < 16; ++attempt) {
    const uint4 x = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), j, attempt), key);
    const uint32_t cand = __umulhi(x.x, I);
    if (cand != p) { item = cand; break; }
  }
  out[t] = (int32_t)item;
}

// The same draws, written as the expanded 1:n_neg BPR batch the step consumes: triple t = (sample s, negative j) ->
// users[t] = user[s], items[2t] = positive[s], items[2t+1] = negative (one 8-byte store) -- no host-side expand.
__global__ void __launch_bounds__(256)
sample_bpr_batch_kernel(const int32_t* __restrict__ users, const int32_t* __restrict__ pos, long long n, int n_neg, uint32_t I,
                        uint2 key, unsigned long long offset, int32_t* __restrict__ out_users, int2* __restrict__ out_items) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * n_neg) return;
  const long long s = t / n_neg;
  const uint32_t j = (uint32_t)(t % n_neg);
  const unsigned long long g = offset + (unsigned long long)s;
  const uint32_t p = (uint32_t)pos[s];
  uint32_t item = (p + 1u) % I;
  for (uint32_t attempt = 0; attempt < 16; ++attempt) {
    const uint4 x = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), j, attempt), key);
    const uint32_t cand = __umulhi(x.x, I);
    if (cand != p) { item = cand; break; }
  }
  out_users[t] = users[s];
  out_items[t] = make_int2((int32_t)p, (int32_t)item);
}

// raw generator output, exported so the tests can pin it to the published known-answer vectors
__global__ void philox_kat_kernel(const uint32_t* __restrict__ in, int n, uint32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4 x = philox4x32_10(make_uint4(in[6 * i], in[6 * i + 1], in[6 * i + 2], in[6 * i + 3]),
                                make_uint2(in[6 * i + 4], in[6 * i + 5]));
  out[4 * i] = x.x; out[4 * i + 1] = x.y; out[4 * i + 2] = x.z; out[4 * i + 3] = x.w;
}

}  // namespace fr

extern "C" int fr_sample_negatives(fr_handle h, const int32_t* pos_items, int64_t n, int32_t n_neg, uint64_t seed,
                                   uint64_t sample_offset, int32_t* out, fr_stream s) {
  if (!h || !pos_items || !out || n < 0 || n_neg <= 0) return FR_ERR_ARG;
  if (h->cfg.num_items < 2) return fail(h, FR_ERR_ARG, "negative sampling needs at least 2 recipes");
  if (n == 0) return FR_OK;
  const long long total = (long long)n * n_neg;
  sample_negatives_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(s)>>>(
      pos_items, (long long)n, n_neg, (uint32_t)h->cfg.num_items, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)),
      (unsigned long long)sample_offset, out);
  ++g_launches;
  FR_CHECK_LAUNCH(h);
  return FR_OK;
}

extern "C" int fr_sample_bpr_batch(fr_handle h, const int32_t* users, const int32_t* pos_items, int64_t n, int32_t n_neg,
                                   uint64_t seed, uint64_t sample_offset, int64_t num_items, int32_t* out_users,
                                   int32_t* out_items, fr_stream s) {
  if (!h || !users || !pos_items || !out_users || !out_items || n < 0 || n_neg <= 0) return FR_ERR_ARG;
  if (num_items == 0) num_items = h->cfg.num_items;
  if (num_items < 2 || num_items > 0x7fffffffll) return fail(h, FR_ERR_ARG, "negative sampling needs 2 <= num_items < 2^31");
  if ((reinterpret_cast<uintptr_t>(out_items) & 7) != 0) return fail(h, FR_ERR_ARG, "out_items must be 8-byte aligned");
  if (n == 0) return FR_OK;
  const long long total = (long long)n * n_neg;
  sample_bpr_batch_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(s)>>>(
      users, pos_items, (long long)n, n_neg, (uint32_t)num_items, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)),
      (unsigned long long)sample_offset, out_users, reinterpret_cast<int2*>(out_items));
  ++g_launches;
  FR_CHECK_LAUNCH(h);
  return FR_OK;
}

extern "C" int fr_philox4x32_10(fr_handle h, const uint32_t* ctr_key /* [n,6] device */, int32_t n, uint32_t* out /* [n,4] */,
                                fr_stream s) {
  if (!h || !ctr_key || !out || n < 0) return FR_ERR_ARG;
  if (n == 0) return FR_OK;
  philox_kat_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(s)>>>(ctr_key, n, out);
  ++g_launches;
  FR_CHECK_LAUNCH(h);
  return FR_OK;
}

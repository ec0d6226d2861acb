// Shared device helpers for the foodrec_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define FR_FULL 0xffffffffu
#define FR_WARPS_PER_BLOCK 8
#define FR_THREADS (FR_WARPS_PER_BLOCK * 32)

namespace fr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FR_FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FR_FULL, v, o);
  return v;
}
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}
__device__ __forceinline__ void fma4(float4& acc, const float s, const float4 v) {
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y);
  acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}
// acc += round(s*v): the product is rounded BEFORE the add, like TF's gradient slices
// (IndexedSlices values are materialised, then segment-summed).  With an FMA the exact
// cancellations of the reference (e.g. a BPR pair whose two recipes share a category mask:
// g*a*pc - g*a*pc == 0) would leave a rounding residual that Adam's 1/eps gain amplifies.
__device__ __forceinline__ void mad4_rn(float4& acc, const float s, const float4 v) {
  acc.x = __fadd_rn(acc.x, __fmul_rn(s, v.x)); acc.y = __fadd_rn(acc.y, __fmul_rn(s, v.y));
  acc.z = __fadd_rn(acc.z, __fmul_rn(s, v.z)); acc.w = __fadd_rn(acc.w, __fmul_rn(s, v.w));
}
__device__ __forceinline__ float4 div4_rn(const float4 v, const float n) {
  return make_float4(__fdiv_rn(v.x, n), __fdiv_rn(v.y, n), __fdiv_rn(v.z, n), __fdiv_rn(v.w, n));
}
__device__ __forceinline__ float4 add4(const float4 a, const float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 scale4(const float s, const float4 v) {
  return make_float4(s * v.x, s * v.y, s * v.z, s * v.w);
}
__device__ __forceinline__ float4 shfl4(const float4 v, const int src) {
  return make_float4(__shfl_sync(FR_FULL, v.x, src), __shfl_sync(FR_FULL, v.y, src),
                     __shfl_sync(FR_FULL, v.z, src), __shfl_sync(FR_FULL, v.w, src));
}

// 16-byte row access.  A row of D floats is DV = D/4 float4; lane l owns float4
// l, l+32, ... (NV of them).  Lanes past DV carry zeros.
__device__ __forceinline__ float4 ld_stream(const float4* p) {   // read-once data: keep out of L1
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Fire-and-forget TMA bulk prefetch of a contiguous span into L2 (SASS: UBLKPF).  A row
// mover is latency-bound per warp (its registers hold one table row at a time); prefetching
// the rows of the warp's NEXT work item turns their demand loads into L2 hits and puts
// more bytes in flight without holding registers.  Measured on B200 (bench A/B, B=65536):
// forward 0.226 -> 0.186 ms with it (used there); the segment-reduce passes got 5 % slower
// (they already run at ~80 % of the HBM peak; kept behind -DFR_PREFETCH_SEG).
// -DFR_PREFETCH_LINES swaps the bulk op for per-128 B prefetch.global.L2 (same result).
__device__ __forceinline__ void prefetch_l2_warp(const void* p, uint32_t bytes, int lane) {   // all lanes call
#if defined(FR_PREFETCH_LINES)
  for (uint32_t o = (uint32_t)lane * 128u; o < bytes; o += 32u * 128u)
    asm volatile("prefetch.global.L2 [%0];" :: "l"(reinterpret_cast<const char*>(p) + o));
#else
  if (lane == 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes) : "memory");
#endif
}

template <int NV>
__device__ __forceinline__ void load_row(float4 (&dst)[NV], const float4* row, int DV, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    dst[k] = (i < DV) ? ld_stream(row + i) : f4zero();
  }
}
template <int NV>
__device__ __forceinline__ void load_row_ro(float4 (&dst)[NV], const float4* __restrict__ row, int DV, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    dst[k] = (i < DV) ? __ldg(row + i) : f4zero();
  }
}
template <int NV>     // read-once rows that must not displace an L2-resident table: ld.global.cs (evict-first)
__device__ __forceinline__ void load_row_cs(float4 (&dst)[NV], const float4* row, int DV, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    dst[k] = (i < DV) ? __ldcs(row + i) : f4zero();
  }
}
template <int NV>
__device__ __forceinline__ void store_row(float4* row, const float4 (&src)[NV], int DV, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i < DV) st_stream(row + i, src[k]);
  }
}

// ---- table storage formats (fr_set_table_format).  A Personal_Memory / Recipe_Embedding row is either fp32 (one float4
// per 4 elements) or bf16 (one uint2 per 4 elements: BASELINE configs[4]).  All arithmetic is fp32; a bf16 table is
// converted on load and rounded to nearest-even on store.  Index arithmetic is in 4-element units either way: kernels keep
// their `float4*` table pointers and go through tab_at<BF>() for the address.
template <bool BF> struct TabVec { using type = float4; };
template <> struct TabVec<true> { using type = uint2; };
template <bool BF>
__device__ __forceinline__ const typename TabVec<BF>::type* tab_at(const float4* base, size_t idx4) {
  return reinterpret_cast<const typename TabVec<BF>::type*>(base) + idx4;
}
template <bool BF>
__device__ __forceinline__ typename TabVec<BF>::type* tab_at(float4* base, size_t idx4) {
  return reinterpret_cast<typename TabVec<BF>::type*>(base) + idx4;
}
__device__ __forceinline__ float4 bf4_to_f4(const uint2 v) {
  return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                     __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
}
__device__ __forceinline__ uint2 f4_to_bf4(const float4 v) {           // round to nearest even
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 r;
  r.x = *reinterpret_cast<const uint32_t*>(&a); r.y = *reinterpret_cast<const uint32_t*>(&b);
  return r;
}
__device__ __forceinline__ float4 tab_ld_ro(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float4 tab_ld_ro(const uint2* p) { return bf4_to_f4(__ldg(p)); }
__device__ __forceinline__ float4 tab_ld_cs(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ float4 tab_ld_cs(const uint2* p) { return bf4_to_f4(__ldcs(p)); }
__device__ __forceinline__ float4 tab_ld_stream(const float4* p) { return ld_stream(p); }
__device__ __forceinline__ float4 tab_ld_stream(const uint2* p) { return bf4_to_f4(__ldcs(p)); }
__device__ __forceinline__ void tab_st_cs(float4* p, const float4 v) { __stcs(p, v); }
__device__ __forceinline__ void tab_st_cs(uint2* p, const float4 v) { __stcs(p, f4_to_bf4(v)); }
// 16-byte-vector rows of either format: lane l owns 4-element group l, l+32, ...
template <int NV, class V>
__device__ __forceinline__ void load_row_t(float4 (&dst)[NV], const V* row, int DV, int lane) {       // read once
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    dst[k] = (i < DV) ? tab_ld_stream(row + i) : f4zero();
  }
}
template <int NV, class V>
__device__ __forceinline__ void load_row_cs_t(float4 (&dst)[NV], const V* row, int DV, int lane) {     // evict-first
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    dst[k] = (i < DV) ? tab_ld_cs(row + i) : f4zero();
  }
}
template <int NV, class V>
__device__ __forceinline__ void load_row_ro_t(float4 (&dst)[NV], const V* __restrict__ row, int DV, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    dst[k] = (i < DV) ? tab_ld_ro(row + i) : f4zero();
  }
}

// health term: pr += alpha * (sum_l G[l]) / n over the labels of user u (see internal.h:HealthBlend)
template <int NV, class HB>
__device__ __forceinline__ void health_blend_rows(float4 (&pr)[5][NV], const HB& hb, int u, int DV, int lane) {
  if (!hb.G) return;
  const int b = hb.lab_off[u], e = hb.lab_off[u + 1];
  if (e <= b) return;
  const float n = (float)(e - b);
#pragma unroll
  for (int s = 0; s < 5; ++s)
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = lane + 32 * k;
      if (i < DV) {
        float4 acc = f4zero();
        for (int x = b; x < e; ++x) acc = add4(acc, __ldg(hb.G + ((size_t)hb.lab_idx[x] * 5 + s) * DV + i));
        float4 q = div4_rn(acc, n);
        pr[s][k].x = __fadd_rn(pr[s][k].x, __fmul_rn(hb.alpha, q.x));
        pr[s][k].y = __fadd_rn(pr[s][k].y, __fmul_rn(hb.alpha, q.y));
        pr[s][k].z = __fadd_rn(pr[s][k].z, __fmul_rn(hb.alpha, q.z));
        pr[s][k].w = __fadd_rn(pr[s][k].w, __fmul_rn(hb.alpha, q.w));
      }
    }
}
// the same on a flat [5*D] float row in shared memory (block-wide)
template <class HB>
__device__ __forceinline__ void health_blend_flat(float* sP, const HB& hb, int u, int D, int tid, int nthreads) {
  if (!hb.G) return;
  const int b = hb.lab_off[u], e = hb.lab_off[u + 1];
  if (e <= b) return;
  const float n = (float)(e - b);
  const float* Gf = reinterpret_cast<const float*>(hb.G);
  for (int i = tid; i < 5 * D; i += nthreads) {
    float acc = 0.f;
    for (int x = b; x < e; ++x) acc = __fadd_rn(acc, __ldg(Gf + (size_t)hb.lab_idx[x] * 5 * D + i));
    sP[i] = __fadd_rn(sP[i], __fmul_rn(hb.alpha, __fdiv_rn(acc, n)));
  }
}

// sum_c m_c * Cat[c]  (un-normalised pooled category row; Model_Recommender.py:67,124-128)
template <int NV>
__device__ __forceinline__ void pooled_cat(float4 (&pc)[NV], const float4* sCat, const float4 m, int DV, int lane) {
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i < DV) {
      const float4 c0 = sCat[i], c1 = sCat[DV + i], c2 = sCat[2 * DV + i], c3 = sCat[3 * DV + i];
      float4 r;
      r.x = m.x * c0.x + m.y * c1.x + m.z * c2.x + m.w * c3.x;
      r.y = m.x * c0.y + m.y * c1.y + m.z * c2.y + m.w * c3.y;
      r.z = m.x * c0.z + m.y * c1.z + m.z * c2.z + m.w * c3.z;
      r.w = m.x * c0.w + m.y * c1.w + m.z * c2.w + m.w * c3.w;
      pc[k] = r;
    } else {
      pc[k] = f4zero();
    }
  }
}
__device__ __forceinline__ float4 div4(const float4 v, const float n) {
  return make_float4(v.x / n, v.y / n, v.z / n, v.w / n);
}
__device__ __forceinline__ float comp(const float4 v, const int c) {
  return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
}

}  // namespace fr

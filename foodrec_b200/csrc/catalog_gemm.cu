// catalog_gemm_kernel<CG>: the dense contraction of full-catalog scoring on tcgen05, with the
// running top-K filter fused into the TMEM epilogue (see catalog.cuh for the maths).
//
// One CTA (CG = 1) or one CTA pair (CG = 2, tcgen05 cta_group::2, M = 256) owns a block of
// users for a whole sweep over its split of the recipe tiles:
//   warp 0   TMA producer: B tiles (recipes, bf16, 128B-swizzled) stream through an mbarrier
//            ring; the A block (one mask's user vectors, all of K) is loaded only when the
//            sweep enters a new mask group and then stays in shared memory (A-stationary: the
//            L2->SM traffic is the B stream alone, 1/(128*CG) bytes per flop)
//   warp 1   one thread issues tcgen05.mma (128*CG x 256 x 16 per instruction) into one of two
//            256-column fp32 accumulator stages in TMEM, commits free the smem stage / publish
//            the accumulator
//   warp 2   TMEM allocation
//   warps 4-7 epilogue: thread r owns user row r for the whole sweep (threshold and candidate
//            count live in its registers): tcgen05.ld 32 columns -> max-tree -> one compare;
//            only a chunk that holds a value >= threshold takes the per-value push path.  A
//            row whose candidate list is nearly full is compacted by its warp (exact K-th
//            largest by bitwise binary search with warp REDUX, keep >= kth - 2E).
// No score matrix is ever written: HBM traffic is B once per concurrent wave plus O(K log) candidates.
#include "catalog.cuh"
#include "common.cuh"
#include "internal.h"
#include "tc05.cuh"

namespace fr {

template <int CG>
struct CatCfg {
  static constexpr int B_ROWS = CAT_BN / CG;            // recipe rows this CTA loads per stage
  static constexpr int B_STAGE = B_ROWS * CAT_BK * 2;   // bytes
  static constexpr int NS = CAT_B_TOTAL / B_STAGE;      // 4 (CG=1) / 8 (CG=2)
};

// Warp-collective compaction of lane L's candidate row.  Keeps every entry >= kth - 2E.
__device__ __noinline__ void catalog_warp_compact(const int L, const int lane, float* __restrict__ cand_sc,
                                                  int32_t* __restrict__ cand_row, const size_t my_base, int& cnt,
                                                  float& thr, const float my_margin2, const int K, int32_t* ovf_flag) {
  constexpr int NV = CAT_CAP / 32;
  const size_t base = __shfl_sync(FR_FULL, (unsigned long long)my_base, L);
  const int n = min(__shfl_sync(FR_FULL, cnt, L), CAT_CAP);
  const float m2 = __shfl_sync(FR_FULL, my_margin2, L);
  float sc[NV]; int32_t rw[NV]; uint32_t key[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e = i * 32 + lane;
    const bool ok = e < n;
    sc[i] = ok ? __ldcg(cand_sc + base + e) : 0.f;
    rw[i] = ok ? __ldcg(cand_row + base + e) : -1;
    key[i] = ok ? fkey(sc[i]) : 0u;
  }
  uint32_t res = 0;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t trial = res | (1u << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) c += (key[i] >= trial);
    c = __reduce_add_sync(FR_FULL, c);
    if (c >= K) res = trial;
  }
  const float nthr = __fsub_rd(funkey(res), m2);
  int out = 0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const bool keep = (i * 32 + lane < n) && (sc[i] >= nthr);
    const uint32_t bal = __ballot_sync(FR_FULL, keep);
    if (keep) {
      const int pos = out + __popc(bal & ((1u << lane) - 1u));
      __stcg(cand_sc + base + pos, sc[i]);
      __stcg(cand_row + base + pos, rw[i]);
    }
    out += __popc(bal);
  }
  __syncwarp();
  if (lane == L) {
    cnt = out; thr = nthr;
    if (out > CAT_CAP - 64) { *ovf_flag = 1; thr = __int_as_float(0x7f800000); }   // too many near-ties: exact fallback
  }
}

template <int CG>
__global__ void __launch_bounds__(CAT_THREADS, 1)
catalog_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const CatGemmParams p) {
  using C = CatCfg<CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + CAT_KB_MAX * CAT_A_BLK;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + CAT_B_TOTAL);
  uint64_t* empty = full + C::NS;
  uint64_t* tfull = empty + C::NS;
  uint64_t* tempty = tfull + 2;
  uint64_t* a_empty = tempty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(a_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? tc::cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) { tc::tma_prefetch_desc(&tmA); tc::tma_prefetch_desc(&tmB); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::NS; ++i) { tc::mbar_init(&full[i], CG); tc::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(&tfull[i], 1); tc::mbar_init(&tempty[i], 4 * CG); }
    tc::mbar_init(a_empty, 1);
    tc::fence_barrier_init();
  }
  if (CG == 2) tc::cluster_sync_all();      // both CTAs resident and initialised before the pair allocates
  if (warp == 2) { tc::tmem_alloc<CG>(tmem_ptr, 512); tc::tmem_relinquish<CG>(); }
  tc::fence_before_sync();
  if (CG == 2) tc::cluster_sync_all(); else __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  const int n_units = p.m_blocks * p.n_split;
  const int cluster_id = blockIdx.x / CG, n_clusters = gridDim.x / CG;
  const int kbn = p.k_blocks;

  if (warp == 0 && lane == 0) {
    // ================= TMA producer =================
    uint32_t stage = 0, phase = 0, a_loads = 0;
    for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
      const int sp = unit / p.m_blocks, mb = unit % p.m_blocks;
      const int t0 = sp * p.tiles_per_split, t1 = min(t0 + p.tiles_per_split, p.n_tiles);
      int prev_g = -1;
      for (int t = t0; t < t1; ++t) {
        const int g = __ldg(p.tile_group + t);
        const bool loadA = g != prev_g;
        prev_g = g;
        if (loadA) { tc::mbar_wait(a_empty, (a_loads & 1u) ^ 1u); ++a_loads; }   // MMAs of the previous A are done
        for (int kb = 0; kb < kbn; ++kb) {
          tc::mbar_wait(&empty[stage], phase ^ 1u);
          uint32_t bytes = C::B_STAGE;
          if (loadA && kb == 0) bytes += kbn * CAT_A_BLK;
          if (CG == 1) tc::mbar_arrive_expect_tx(&full[stage], bytes);
          else if (leader) tc::mbar_arrive_expect_tx(&full[stage], 2 * bytes);
          else tc::mbar_arrive_cluster(&full[stage], 0);
          if (loadA && kb == 0) {
            const int arow = g * p.m_pad + mb * (CAT_BM * CG) + rank * CAT_BM;
            for (int k2 = 0; k2 < kbn; ++k2)
              tc::tma_load_2d<CG>(sA + k2 * CAT_A_BLK, &tmA, &full[stage], k2 * CAT_BK, arow);
          }
          tc::tma_load_2d<CG>(sB + stage * C::B_STAGE, &tmB, &full[stage], kb * CAT_BK,
                              t * CAT_BN + (int)rank * C::B_ROWS);
          if (++stage == C::NS) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ================= MMA issuer (one thread of the leader CTA) =================
    constexpr uint32_t idesc = tc::umma_idesc_bf16(CAT_BM * CG, CAT_BN);
    uint32_t stage = 0, phase = 0, tcount = 0;
    for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
      const int sp = unit / p.m_blocks;
      const int t0 = sp * p.tiles_per_split, t1 = min(t0 + p.tiles_per_split, p.n_tiles);
      for (int t = t0; t < t1; ++t, ++tcount) {
        const uint32_t as = tcount & 1u, aph = (tcount >> 1) & 1u;
        const bool a_last = (t + 1 == t1) || (__ldg(p.tile_group + t + 1) != __ldg(p.tile_group + t));
        tc::mbar_wait(&tempty[as], aph ^ 1u);          // epilogue drained this accumulator stage
        tc::fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * CAT_BN;
        for (int kb = 0; kb < kbn; ++kb) {
          tc::mbar_wait(&full[stage], phase);
          tc::fence_after_sync();
          const uint64_t ad = tc::umma_desc_sw128(sA + kb * CAT_A_BLK);
          const uint64_t bd = tc::umma_desc_sw128(sB + stage * C::B_STAGE);
#pragma unroll
          for (int k = 0; k < CAT_BK / 16; ++k)      // +32 B per k16 step inside the swizzle atom (addr field is >>4)
            tc::mma_bf16<CG>(d_tmem, ad + 2u * k, bd + 2u * k, idesc, (kb | k) ? 1u : 0u);
          tc::mma_commit<CG>(&empty[stage]);
          if (kb == kbn - 1) {
            tc::mma_commit<CG>(&tfull[as]);
            if (a_last) tc::mma_commit<CG>(a_empty);
          }
          if (++stage == C::NS) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: running top-K filter =================
    const int q = warp & 3;
    const int r_in_blk = q * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float INF = __int_as_float(0x7f800000);
    uint32_t tcount = 0;
    for (int unit = cluster_id; unit < n_units; unit += n_clusters) {
      const int sp = unit / p.m_blocks, mb = unit % p.m_blocks;
      const int t0 = sp * p.tiles_per_split, t1 = min(t0 + p.tiles_per_split, p.n_tiles);
      const int grow = mb * (CAT_BM * CG) + (int)rank * CAT_BM + r_in_blk;
      const bool valid = grow < p.n_rows;
      const size_t base = (static_cast<size_t>(sp) * p.m_pad + grow) * CAT_CAP;
      const float m2 = __ldg(p.margin2 + grow);
      float thr = valid ? -INF : INF;
      int cnt = 0;
      int cur_g = -1;
      float bias = 0.f;
      for (int t = t0; t < t1; ++t, ++tcount) {
        const uint32_t as = tcount & 1u, aph = (tcount >> 1) & 1u;
        const int g = __ldg(p.tile_group + t);
        if (g != cur_g) { cur_g = g; bias = __ldg(p.bias + (size_t)g * p.m_pad + grow); }
        float adj = __fsub_rd(thr, bias);            // push iff  v + bias >= thr
        tc::mbar_wait(&tfull[as], aph);
        tc::fence_after_sync();
        const int n0 = t * CAT_BN;
        const int nvalid = __ldg(p.tile_valid + t);     // rows of this tile that hold a recipe (the rest is zero padding)
#pragma unroll 1
        for (int c = 0; c < CAT_BN / 32; ++c) {
          float v[32];
          __syncwarp();
          tc::tmem_ld_32x32(t_lane + as * CAT_BN + c * 32, v);
          tc::tmem_ld_wait();
          if (c == CAT_BN / 32 - 1) {                 // accumulator fully read: hand it back to the MMA warp
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) { if (CG == 1) tc::mbar_arrive(&tempty[as]); else tc::mbar_arrive_cluster(&tempty[as], 0); }
          }
          float m16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) m16[j] = fmaxf(v[2 * j], v[2 * j + 1]);
#pragma unroll
          for (int j = 0; j < 8; ++j) m16[j] = fmaxf(m16[2 * j], m16[2 * j + 1]);
          const float mx = fmaxf(fmaxf(fmaxf(m16[0], m16[1]), fmaxf(m16[2], m16[3])),
                                 fmaxf(fmaxf(m16[4], m16[5]), fmaxf(m16[6], m16[7])));
          if (mx >= adj) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (v[j] >= adj && c * 32 + j < nvalid) {
                const int prow = n0 + c * 32 + j;
                if (cnt < CAT_CAP) { __stcg(p.cand_sc + base + cnt, v[j] + bias); __stcg(p.cand_row + base + cnt, prow); }
                ++cnt;
              }
            }
          }
          uint32_t need = __ballot_sync(FR_FULL, cnt > CAT_CAP - 32);
          while (need) {
            const int L = __ffs(need) - 1;
            need &= need - 1;
            catalog_warp_compact(L, lane, p.cand_sc, p.cand_row, base, cnt, thr, m2, p.K, p.ovf + grow);
            if (lane == L) adj = __fsub_rd(thr, bias);
          }
        }
      }
      if (valid) p.cand_cnt[(size_t)sp * p.m_pad + grow] = min(cnt, CAT_CAP);
    }
  }

  __syncwarp();
  tc::fence_before_sync();
  if (CG == 2) tc::cluster_sync_all(); else __syncthreads();
  if (warp == 2) tc::tmem_dealloc<CG>(tmem_base, 512);
}

cudaError_t catalog_gemm_configure() {
  cudaError_t e = cudaFuncSetAttribute(catalog_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CAT_SMEM);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(catalog_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CAT_SMEM);
}

void launch_catalog_gemm(int cta_group, int sm_count, const CUtensorMap& tmA, const CUtensorMap& tmB,
                         const CatGemmParams& p, cudaStream_t st) {
  const int n_units = p.m_blocks * p.n_split;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(CAT_THREADS);
  cfg.dynamicSmemBytes = CAT_SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (cta_group == 2) {
    const int pairs = min(sm_count / 2, n_units);
    cfg.gridDim = dim3(2 * (pairs < 1 ? 1 : pairs));
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, catalog_gemm_kernel<2>, tmA, tmB, p);
  } else {
    const int ctas = min(sm_count, n_units);
    cfg.gridDim = dim3(ctas < 1 ? 1 : ctas);
    cfg.attrs = nullptr; cfg.numAttrs = 0;
    cudaLaunchKernelEx(&cfg, catalog_gemm_kernel<1>, tmA, tmB, p);
  }
  ++g_launches;
}

}  // namespace fr

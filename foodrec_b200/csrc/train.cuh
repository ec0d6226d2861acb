// Parameter blocks for the train-step kernels (train.cu) and their launchers.
#pragma once
#include "internal.h"

namespace fr {

struct FwdParams {
  const float4 *P, *R, *cat;        // cat = step-start snapshot of Category_Embedding
  int DV, B;
  float Bnorm;                      // batch size the loss mean divides by (GLOBAL batch when sharded)
  const int32_t *users, *items;
  const float4* cats; int cats_by_item;
  const float* labels;
  float a, oma;
  float* g; float4* z; float* scores;
  float *part_loss, *part_nrm; float4* part_gcat;   // one slot per block
  // lazy-exact Adam: P[u] in memory may be stale; the forward replays the skipped decay
  // steps last+1..step-1 in registers from (m, v) so that it scores the row TF would see
  int lazy; const float4 *mP, *vP; const int32_t* lastP; OptConsts oc;
  int tab;                          // table storage: 0 fp32 P and R, 1 bf16 P and R, 2 bf16 P + fp32 R (fr_set_table_format)
};

struct FinalizeParams {
  const float *part_loss, *part_nrm; const float4* part_gcat; int nblk;
  int DV; float inv_count_B;  // loss mean divisor (global batch)
  float B;
  float* packed;          // [2 + 4*D]: loss_sum, nrm_sum, gCat  (all-reduced across ranks in multi-GPU)
  int do_reduce, do_apply;
  float4 *Cat, *s1Cat, *s2Cat;
  OptConsts oc; float clip;
  float* out; float* lr_hist;
};

struct SegCommon {
  const uint32_t* keys; const uint32_t* perm;
  const uint32_t* n_dev; uint32_t n_host;
  float4* pieces;              // [(chunk*2+slot) * NR*DV]
  uint32_t* uniq_counter;      // nullable
  // Long crossing runs (a hot recipe, every label): the per-warp combine hands a chain of more than
  // FR_LONG_CHAIN pieces to seg_combine_long_kernel (one block per chain) through this list.  Nullable.
  uint4* long_list;            // {first piece index, its slot, last piece index, key}
  uint32_t* long_count; uint32_t long_cap;
  const float* only_if_scaled; // nullable: out scalars; the pass exits at once if out[FR_OUT_SCALE] == 1 (the update pass
                               // that follows a single-pass step whose clip speculation held has nothing to do)
};
constexpr uint32_t FR_LONG_CHAIN = 16;

struct UserPolParams {
  float4 *P, *s1, *s2; int32_t* last;
  const float4 *R, *G; const float4* cat;   // pre-step values
  const int32_t* items; const float* g; const float4* cats; int cats_by_item;
  const float* ws_row; const float* out;    // out[FR_OUT_SCALE]
  int group; ModelConsts mc; OptConsts oc;
  const float* user_labels; const int32_t *lab_off, *lab_idx; const int32_t* users;
  int tab;                                  // as FwdParams::tab
};
// single-pass step (user_fused_kernel, train_seg.cu): forward + segment reduce + Adam in one walk over the user-sorted rows
struct FusedParams {
  float4 *P[2], *m[2], *v[2];        // [0] the caller's tables, [1] the shadow copies (fr_set_shadow)
  int32_t* last;                      // bit 30: which copy holds the row; low bits: lazy-Adam stamp
  const float4 *R, *cat;
  const int32_t* items; const float4* cats; int cats_by_item;
  const float* labels;                // pointwise: y of each row
  float a, oma, Bnorm;
  float* g; float4* z; float* scores;
  float *part_loss, *part_nrm; float4* part_gcat;     // one slot per block
  ModelConsts mc; OptConsts oc;
};
int user_fused_grid(uint32_t n_rows, int sm_count);
void launch_user_fused(int NV, int group, const SegCommon& c, const FusedParams& p, int grid, const Launch& l);
void launch_user_commit(const uint32_t* keys, uint32_t n, int32_t* last, const float* out, int step, const Launch& l);
void launch_shadow_consolidate(int32_t* last, int64_t n_users, int rowDV, float4* P, float4* m, float4* v, const float4* Pa,
                               const float4* ma, const float4* va, const float* pred, const Launch& l);

struct ItemPolParams {
  float4 *R, *s1, *s2; int32_t* last;
  const float4* z; const float* g; const float* out;
  ModelConsts mc; OptConsts oc;
  int tab;                                  // != 0: Recipe_Embedding is stored in bf16
};
struct LabelPolParams {
  float4* G; const float4 *R, *cat;
  const uint32_t* ent_row; const float* ent_coef;
  const int32_t* items; const float4* cats; int cats_by_item;
  ModelConsts mc;
  int tab;                                  // 1: the recipe rows read here are bf16
};

struct LabelEmitParams {
  int S, group, L; const int32_t* users;
  const float* user_labels; const int32_t *lab_off, *lab_idx;
  const float* ws_row;
  uint32_t* counts;     // [S]
  const uint32_t* offs; // [S]
  uint32_t *ent_key, *ent_row; float* ent_coef; uint32_t cap;
  uint32_t* n_entries; float* out;
};

void launch_prep_rows(int mode, int B, const int32_t* users, const int32_t* items, const float* labels, const float* ws_in,
                      int64_t n_users, int64_t n_items, uint32_t* ukeys, float* ws_row, int32_t* users_s, int32_t* items_s,
                      float* flag, const Launch& l);
void launch_fwd_train(int NV, int group, const FwdParams& p, int grid, const Launch& l);
int fwd_train_grid(int B, int sm_count);
void launch_finalize(const FinalizeParams& p, const Launch& l);
void launch_user_pass(int NV, const SegCommon& c, const UserPolParams& p, const Launch& l);
void launch_personal_pass(int NV, const SegCommon& c, const UserPolParams& p, const Launch& l);
void launch_item_pass(int NV, const SegCommon& c, const ItemPolParams& p, const Launch& l);
void launch_label_pass(int NV, const SegCommon& c, const LabelPolParams& p, const Launch& l);
void launch_label_count(const LabelEmitParams& p, const Launch& l);
void launch_label_emit(const LabelEmitParams& p, const Launch& l);
void launch_adam_sweep(float4* var, float4* m, float4* v, int32_t* last, int64_t nrows, int rowDV,
                       const OptConsts& oc, int target_step, const Launch& l);
void launch_fill_i32(int32_t* p, int64_t n, int32_t v, const Launch& l);
void launch_series_update(double* cser, const float* lr_hist, int t, float b1, float b2, const Launch& l);
void launch_series_rebuild(double* cser, const float* lr_hist, int step, float b1, float b2, const Launch& l);
void launch_item_catchup(int NV, const uint32_t* keys, uint32_t n, float4* R, float4* m, float4* v,
                         int32_t* last, int DV, const OptConsts& oc, const Launch& l,
                         const uint32_t* n_dev = nullptr);
// Peer-memory exchange (NVLink P2P): instead of staging rows for an all-to-all, the producing kernel stores each
// row straight into the consumer rank's buffer.  dst[w] = rank w's buffer [W*cap, DV]; a row for rank w's slot j goes
// to dst[w] + (my_rank*cap + j)*DV -- exactly where the all-to-all would have put it.  world == 0: not in use.
struct PeerPtrs { float4* dst[8]; int world, rank, cap; };
void launch_item_grad_pass(int NV, const SegCommon& c, const ItemPolParams& p, float4* gbuf, const PeerPtrs& peers, const Launch& l);

// ---- row-sharded training helpers (train_shard.cu)
struct ShardPlanParams {
  int S, W, cap; uint32_t items_per_rank;
  const int32_t* items;          // [S] GLOBAL recipe ids
  const float4* item_cats;       // [I] replicated, or null when per-row cats are fed
  const float4* cats_in;         // [S] per-row cats (or null)
  uint32_t* okeys;               // [S] owner-major key  (id % W) * items_per_rank + id / W
  float4* cats_row;              // [S] gathered per-row categories
  // after the sort by okey:
  const uint32_t* okeys_sorted; const uint32_t* perm;
  uint32_t* flags; uint32_t* excl; uint32_t* owner_counts;   // [S], [S], [W]
  int32_t* req;                  // [W*cap] local row at the owner, -1 = empty
  int32_t* slot_of_row;          // [S]
  uint32_t* slot_sorted;         // [S] slot of sorted position (monotone: doubles as sort key)
  float* out;                    // FR_OUT_OVERFLOW on capacity overflow
};
void launch_shard_prep(const ShardPlanParams& p, const Launch& l);
void launch_shard_heads(const ShardPlanParams& p, const Launch& l);
void launch_shard_fill(const ShardPlanParams& p, const Launch& l);
void launch_serve_keys(const int32_t* rreq, uint32_t n, uint32_t items_per_rank, uint32_t* keys, uint32_t* n_valid,
                       const Launch& l);
void launch_gather_rows(const float4* R, const int32_t* rreq, uint32_t n, int DV, float4* out, const PeerPtrs& peers, const Launch& l,
                        uint32_t n_table, int bf16);
void launch_add_inplace(float4* dst, const float4* src, int64_t n4, const Launch& l);
void launch_route_keys(const int32_t* users, int B, int W, uint32_t* keys, uint32_t* owner_counts, const Launch& l);
void launch_route_fill(const int32_t* users, const int32_t* items, const float* labels, int B, int W, int group, int rcap,
                       const uint32_t* okeys_sorted, const uint32_t* perm, const uint32_t* owner_counts, int32_t* send,
                       float* flag, const Launch& l);
void launch_route_unpack(const int32_t* recv, int W, int blk, int rcap, int group, int has_labels, uint32_t* counts, int cap_out,
                         int32_t* users, int32_t* items, float* labels, int32_t* n_out, float* flag, const Launch& l);
void launch_mean(const float4* x, int64_t n4, double* partials, float* out_slot, double count, const Launch& l);
void launch_write_counters(const uint32_t* counters, float* out, const Launch& l);

}  // namespace fr

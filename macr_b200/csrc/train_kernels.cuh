// train_kernels.cuh -- declarations shared by train_kernels.cu (kernels + launchers) and
// trainer.cu (the step handles).  Internal; the public surface is include/macr_b200.h.
#pragma once
#include "common.cuh"

namespace macr {

// Device-resident per-trainer state.  One tiny kernel advances it at the start of every step so
// a captured CUDA graph can be replayed without touching kernel parameters.
struct StepState {
  float b1p, b2p;  // beta1^t, beta2^t as TF keeps them (fp32, repeated multiplication)
  float lr_t;      // lr * sqrt(1-b2p) / (1-b1p) for the step being executed
  float pad0;
  const int32_t *ids_base;  // [n_steps][3][B]  users | pos | neg
  float *loss_base;         // [n_steps][4]
  long long step_idx;       // index into ids_base / loss_base for the step being executed
  long long t;              // Adam steps applied so far
  const int32_t *gids_base; // row-partitioned MF: the GLOBAL ids [n_steps][3][B]; ids_base then
                            // points at the renumbered (local) copy the exchange kernel fills
  float *cur_loss;          // loss_base + step_idx*4
};

struct GridWs {  // carve-up of the grid kernel's partial-sum workspace
  int tile_i, tile_j;  // rows x columns of one CTA tile
  int nblk_i, nblk_j;  // ceil(B / tile_i), ceil(B / tile_j)
  int ngrp_j, grp_tiles;  // a CTA walks grp_tiles consecutive column tiles of one row band:
                          // ngrp_j = ceil(nblk_j / grp_tiles) CTAs per band, sized so that the whole
                          // grid is resident at once (row sums stay in registers, one prologue per CTA)
  int Bpad;            // B rounded up to the larger tile
  float *rowP, *rowN;  // [ngrp_j][Bpad]
  float *colP, *colN;  // [nblk_i][Bpad]
  float *losspart;     // [nblk_i*ngrp_j]
  float *litem, *luser;  // [Bpad] branch losses per position (model.py:213,215)
  float *gA, *gAN, *gG;  // [Bpad] sig(sp), sig(sn), sig(su)
  int item_gate_only;    // `--train rubibce` / `--loss bce1` (model.py:158-183): gather_dots writes
                         // gG = 1, luser = 0, so the grid is yp_j*sig(sp_i), L_user and d_su vanish
  size_t part_bytes;     // leading bytes of the workspace that must be 0xff before the first launch
  size_t bytes;
};
GridWs grid_ws_layout(int B, void *base);

struct PlanBufs {  // output of one batch_plan over n_ids ids (+ its scratch)
  int32_t *uniq_rows;  // [n_uniq] table row of every segment, ascending
  int32_t *seg_off;    // [n_uniq+1] offsets into the sorted order
  int32_t *seg_pos;    // [n_ids] batch positions sorted by (row, position)
  int32_t *n_uniq;     // device scalar
  int4 *rec;           // [n_ids] per sorted index: {rank << 16 | slot, row, position, first position
                       //          of the segment}
  int32_t *done;       // [n_ids] per-segment arrival counters (multi-unit segments)
};
PlanBufs plan_carve(int32_t *uniq_rows, int32_t *seg_off, int32_t *seg_pos, int32_t *n_uniq,
                    void *ws, int n_ids);

// launchers (all asynchronous on `s`); *_st variants read ids / lr_t through StepState
int launch_step_state(StepState *st, int B, float lr, float b1, float b2, int train,
                      cudaStream_t s);
int launch_gather_dots(const float *Ue, const float *Ie, const float *Ur, const float *Ir,
                       const float *w, const float *wu, const int32_t *u, const int32_t *p,
                       const int32_t *n, const StepState *st, int B, float *yp, float *yn,
                       float *sp, float *sn, float *su, float *regsq, float *snap,
                       const GridWs *gates, cudaStream_t s);
// gates of the batch into ws (the stand-alone grid entry point; trainers use gather_dots)
int launch_gates(const float *sp, const float *sn, const float *su, int B, const GridWs &ws,
                 cudaStream_t s);
// B x B grid + band folds + the step's loss scalars.  st != nullptr: {loss, mf, reg, L_ori} go
// to st->loss_base + 4*st->step_idx (regsq required); else {L_ori, L_item, L_user} to losses3.
int launch_grid_bce(const float *yp, const float *yn, int B, const macr_hparams &hp,
                    const GridWs &ws, float *d_yp, float *d_yn, float *d_sp, float *d_sn,
                    float *d_su, int want_grad, const float *regsq, const StepState *st,
                    float *losses3, cudaStream_t s, bool pdl = false);
// `--train normalbce`: element-wise BCE on (yp, yn) instead of the B x B grid
int launch_plain_bce(const float *yp, const float *yn, int B, const macr_hparams &hp,
                     const float *regsq, const StepState *st, float *losses_direct, float *d_yp,
                     float *d_yn, float *d_sp, float *d_sn, float *d_su, cudaStream_t s);
size_t plan_ws_bytes(int n_ids);
int plan_init();
// two tables in one launch (table 1 optional: n_ids1 == 0)
int launch_batch_plan2(const int32_t *ids0, const StepState *st0, int ids0_off, int n_ids0,
                       int64_t rows0, PlanBufs out0, uint32_t *bitmap0, const int32_t *ids1,
                       int ids1_off, int n_ids1, int64_t rows1, PlanBufs out1, uint32_t *bitmap1,
                       cudaStream_t s);
int launch_mark_touched(const StepState *st, const int32_t *ids, int B, uint32_t *bmU,
                        uint32_t *bmI, cudaStream_t s);
int launch_adam_sweep2(float *var0, float *m0, float *v0, int64_t rows0, const uint32_t *bm0,
                       float *var1, float *m1, float *v1, int64_t rows1, const uint32_t *bm1,
                       float lr_t, const StepState *st, float b1, float b2, float eps,
                       cudaStream_t s);
// fused Adam on the touched rows (MF) and fused step tail, both optional
struct AdamTabs {  // U == nullptr: no fused Adam, summed rows go to gU / gI (LightGCN)
  float *U, *mU, *vU, *I, *mI, *vI;
  uint32_t *bmU, *bmI;
  float b1, b2, eps, lr;  // lr: lr_t itself when st == nullptr
  const StepState *st;
};
struct TailArgs {
  int fused;  // 1: the last CTA runs the step tail
  float *w, *mw, *vw, *wu, *mwu, *vwu;
  macr_hparams hp;
  StepState *st;
  unsigned *ticket;
  int frozen;  // bit 0: w, bit 1: w_user receive no gradient in this graph -> left untouched
};
// summed gradient rows of the unique touched rows (MF: + L2 term, + Adam on those rows) and
// per-CTA partials of grad(w), grad(w_user); rows come from the gather_dots snapshot
int launch_row_grads(const float *snap, const float *w, const float *wu, int B, const float *d_yp,
                     const float *d_yn, const float *d_sp, const float *d_sn, const float *d_su,
                     float lam, PlanBufs planU, PlanBufs planI, float *gU, float *gI,
                     float *unit_part, float *gw_part, float *gwu_part, int *n_part,
                     const AdamTabs *tabs, const TailArgs *tail, cudaStream_t s, bool pdl = false);
int launch_adam_rows2(float *U, float *mU, float *vU, PlanBufs planU, const float *gU,
                      uint32_t *bmU, float *I, float *mI, float *vI, PlanBufs planI,
                      const float *gI, uint32_t *bmI, int max_rows, float lr_t,
                      const StepState *st, float b1, float b2, float eps, cudaStream_t s);
int launch_adam_vec2(float *w, float *mw, float *vw, float *wu, float *mwu, float *vwu,
                     const float *gw_part, const float *gwu_part, int n_part, float lr_t,
                     const StepState *st, float b1, float b2, float eps, cudaStream_t s);
// last kernel of a LightGCN / loss-only step: ApplyAdam on w / w_user (train), state advance
int launch_step_tail(float *w, float *mw, float *vw, float *wu, float *mwu, float *vwu,
                     const float *gw_part, const float *gwu_part, int n_part,
                     const macr_hparams &hp, StepState *st, int train, cudaStream_t s);
int launch_adam_dense(float *var, float *m, float *v, const float *grad, int64_t n_elems,
                      float lr_t, const StepState *st, float b1, float b2, float eps,
                      cudaStream_t s);
int row_grads_max_parts(int B);

}  // namespace macr

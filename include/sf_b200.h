/* sf_b200.h — C ABI of the B200-native speaker/follower recurrent hot path.
 *
 * The reference (ronghanghu/speaker_follower) has no FFI layer: its hot path is the Python module API
 * of tasks/R2R/model.py.  This header is the boundary a maintainer binds instead of
 * those nn.Module forwards: plain device pointers, sizes and a cudaStream_t passed as void*.  No torch
 * types, no allocation inside the library (the caller provides a workspace), no host synchronisation,
 * safe under CUDA-graph capture.  Every entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - all tensors fp32 row-major contiguous unless a leading dimension is given; indices int32;
 *     masks uint8 (1 = masked / padded), exactly the reference's ByteTensor masks (follower.py:101).
 *   - weights are passed in the reference's own state_dict layouts ([out_features, in_features]
 *     row-major), so nn.Parameter storage is used in place; the only derived copy is the caller-owned
 *     packed blob of sfb_follower_pack_weights (explicit, refreshed by the caller when weights change).
 *   - return value: 0 on success, negative sfb_status on error; sfb_last_error() gives the message of
 *     the last failure on the calling thread.  Argument errors are detected before any launch.
 *   - every function only enqueues work on `stream`.
 */
#ifndef SF_B200_H_
#define SF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB_ABI_VERSION 1

typedef enum sfb_status {
  SFB_OK = 0,
  SFB_ERR_INVALID_ARG = -1,   /* NULL pointer, bad size, unsupported dimension */
  SFB_ERR_WORKSPACE = -2,     /* workspace too small / misaligned */
  SFB_ERR_CUDA = -3,          /* a CUDA runtime call failed (message has the CUDA error string) */
  SFB_ERR_NO_DEVICE = -4      /* no sm_100 device / kernel image not loadable */
} sfb_status;

/* Model dimensions.  Reference values (tasks/R2R/train.py:30-38, model.py:303,335):
 * E = action embedding = 2176, F = visual feature = 2176, H = 512, D = 256, V = 36 views. */
typedef struct sfb_dims {
  int32_t E, F, H, D, V;
} sfb_dims;

/* Visual-attention + LSTMCell weights shared by AttnDecoderLSTM (model.py:371-373) and
 * SpeakerEncoderLSTM (model.py:415-418). */
typedef struct sfb_vis_lstm_weights {
  const float* lstm_w_ih;   /* [4H, E+F]  lstm.weight_ih  (gate order i,f,g,o) */
  const float* lstm_w_hh;   /* [4H, H]    lstm.weight_hh */
  const float* lstm_b_ih;   /* [4H]       lstm.bias_ih */
  const float* lstm_b_hh;   /* [4H]       lstm.bias_hh */
  const float* va_w_h;      /* [D, H]     visual_attention_layer.linear_in_h.weight */
  const float* va_b_h;      /* [D] */
  const float* va_w_v;      /* [D, F]     visual_attention_layer.linear_in_v.weight */
  const float* va_b_v;      /* [D]        (cancels inside the softmax; accepted for completeness) */
} sfb_vis_lstm_weights;

/* SoftDotAttention weights (model.py:117,119). */
typedef struct sfb_softdot_weights {
  const float* w_in;        /* [H, H]   linear_in.weight  (no bias) */
  const float* w_out;       /* [H, 2H]  linear_out.weight (no bias), input order [weighted_ctx ; h] */
} sfb_softdot_weights;

/* EltwiseProdScoring weights (model.py:338-340). */
typedef struct sfb_scoring_weights {
  const float* w_h;         /* [D, H] linear_in_h.weight */
  const float* b_h;         /* [D] */
  const float* w_a;         /* [D, E] linear_in_a.weight */
  const float* b_a;         /* [D] */
  const float* w_out;       /* [1, D] linear_out.weight */
  const float* b_out;       /* [1] */
} sfb_scoring_weights;

/* Where a step's 36-view feature slab comes from (replaces Seq2SeqAgent._feature_variables,
 * follower.py:291-298 + ImageFeatures.batch_features, env.py:330-332).
 *   dense : `visual` = [B, V, F] already on the device (what the reference builds on the host per step);
 *   gather: `visual` = NULL; row v of batch element b is the concatenation of
 *           feat_table[vp_idx[b], v, 0:img_dim] and loc_table[view_idx[b], v, 0:F-img_dim]
 *           (env.py:773: feature = concat(image feature, _static_loc_embeddings[viewIndex])). */
typedef struct sfb_visual_source {
  const float*   visual;      /* [B, V, F] or NULL */
  const float*   feat_table;  /* [n_viewpoints, V, img_dim] */
  const float*   loc_table;   /* [V, V, F - img_dim]  (env.py:100-101) */
  const int32_t* vp_idx;      /* [B] row into feat_table */
  const int32_t* view_idx;    /* [B] agent viewIndex 0..V-1 */
  int32_t        img_dim;     /* 2048 */
  int32_t        idx_dependent; /* 1: vp_idx / view_idx are WRITTEN by work enqueued on the same stream just before this
                                   call (e.g. sfb_nav_step): kernels read them only after their dependency wait instead of
                                   prefetching slabs while the predecessor is still running.  0: they are step inputs
                                   that were complete before the previous kernel started (host copies, earlier steps). */
} sfb_visual_source;

/* ---------------------------------------------------------------------------------------------- */

int32_t     sfb_abi_version(void);
const char* sfb_last_error(void);

/* Runtime options (testing / bring-up hooks; process-wide, not thread-safe against concurrent launches):
 *   "disable_tc" = 1 : run the LSTM-gate GEMM on the exact-fp32 FFMA path instead of tcgen05 (bf16x3);
 *   "tc_debug"       : descriptor-encoding variants of the tensor-core path (bring-up only). */
int32_t sfb_set_option(const char* name, int32_t value);
/* Bring-up: phase timestamps (SM clock) of the tensor-core GEMM's CTA 0, valid after a call with tc_debug bit 2. */
int32_t sfb_debug_read_timestamps(int64_t* out, int32_t n);
/* Bring-up: after sfb_set_option("trace", 1), every kernel enqueued records {entry, after dependency wait, exit}
 * (globaltimer ns, block 0) in launch order; returns the number of slots copied (16 int64 each: 0-2 as above, 3-15 kernel-specific phases). */
int32_t sfb_debug_read_trace(int64_t* out, int32_t max_slots);
/* Bring-up: after sfb_set_option("cta_trace", 1), the attention kernel records per CTA {entry, first row landed,
 * stream done, exit, smid, query ready} (8 int64 per CTA, globaltimer ns); returns the number of CTA records copied. */
int32_t sfb_debug_read_cta_trace(int64_t* out, int32_t max_ctas);

/* Bring-up: how many thread-block clusters of `cluster` CTAs (320 threads, `smem` dynamic bytes each) the device can
 * hold at once (cudaOccupancyMaxActiveClusters on the tensor-core projection kernel); negative on error. */
int32_t sfb_debug_max_active_clusters(int32_t cluster, int32_t smem);

/* Device properties the library was built for / sees.  Fills sm (e.g. 100), number of SMs and max
 * opt-in shared memory per block; returns SFB_ERR_NO_DEVICE when there is no usable device. */
int32_t sfb_device_info(int32_t* sm, int32_t* num_sms, int32_t* smem_per_block);

/* Bytes of device workspace one follower decode step needs (also an upper bound for the speaker
 * encoder step and the two attention entry points at the same B).  256-byte aligned pointer required.
 * WORKSPACE CONTRACT (all *_workspace_bytes functions): the buffer must be zero-filled once before its first
 * use (its head holds self-resetting inter-CTA semaphores); afterwards it can be re-used by any number of
 * calls, but not by two calls that may run concurrently on different streams. */
size_t sfb_follower_step_workspace_bytes(const sfb_dims* dims, int32_t B, int32_t L, int32_t A);

/* VisualSoftDotAttention.forward — model.py:310-326.
 * h [B,H] -> feature [B,F], alpha_v [B,V].  Visual rows are read from HBM exactly once. */
int32_t sfb_visual_attention_fwd(const sfb_dims* dims, const sfb_vis_lstm_weights* w, int32_t B,
                                 const float* h, const sfb_visual_source* vis,
                                 float* feature, float* alpha_v,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* The attention-gather kernel alone (model.py:320-325): given the projected query q [B,F]
 * (= W_v^T (W_h h + b_h)), alpha_v = softmax_v(V_v . q), feature = sum_v alpha_v V_v.  One launch; this is
 * the kernel whose HBM roofline fraction BASELINE.json's "attn HBM %" refers to. */
int32_t sfb_visual_attention_core_fwd(const sfb_dims* dims, int32_t B, const float* q,
                                      const sfb_visual_source* vis, float* feature, float* alpha_v,
                                      void* workspace, size_t workspace_bytes, void* stream);

/* SoftDotAttention.forward — model.py:122-143.
 * h [B,H], ctx [B,L,H], mask [B,L] (may be NULL) -> h_tilde [B,H], alpha [B,L]. */
int32_t sfb_soft_dot_attention_fwd(const sfb_dims* dims, const sfb_softdot_weights* w, int32_t B, int32_t L,
                                   const float* h, const float* ctx, const uint8_t* mask,
                                   float* h_tilde, float* alpha,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* EltwiseProdScoring.forward — model.py:342-352 (its `mask` argument is ignored there too):
 *   logit[b,a] = w_out . ((W_h h~_b + b_h) (.) (W_a u_{b,a} + b_a)) + b_out,  h_tilde [B,H], all_u_t [B,A,E] -> logit [B,A].
 * Workspace: sfb_follower_step_workspace_bytes(dims, B, 1, A). */
int32_t sfb_eltwise_prod_scoring_fwd(const sfb_dims* dims, const sfb_scoring_weights* ws, int32_t B, int32_t A,
                                     const float* h_tilde, const float* all_u_t, float* logit,
                                     void* workspace, size_t workspace_bytes, void* stream);

/* AttnDecoderLSTM.forward — model.py:377-397: ONE follower decode step.
 *   u_prev [B,E], all_u_t [B,A,E], vis (dense or gather), h0,c0 [B,H], ctx [B,L,H], ctx_mask [B,L]|NULL
 *   drop_x [B,E+F] / drop_h [B,H]: scaled keep masks of the two nn.Dropout calls (model.py:392,394),
 *   NULL in eval mode.
 *   -> h1,c1 [B,H] (un-dropped), alpha [B,L], logit [B,A] (raw, unmasked), alpha_v [B,V].
 * Intermediates needed by sfb_follower_step_bwd stay in `workspace` (layout private to the library). */
int32_t sfb_follower_step_fwd(const sfb_dims* dims, const sfb_vis_lstm_weights* wl,
                              const sfb_softdot_weights* wt, const sfb_scoring_weights* ws,
                              int32_t B, int32_t L, int32_t A,
                              const float* u_prev, const float* all_u_t, const sfb_visual_source* vis,
                              const float* h0, const float* c0, const float* ctx, const uint8_t* ctx_mask,
                              const float* drop_x, const float* drop_h,
                              float* h1, float* c1, float* alpha, float* logit, float* alpha_v,
                              void* workspace, size_t workspace_bytes, void* stream);

/* ---- packed-weight fast path of the same step (AttnDecoderLSTM.forward, model.py:377-397) ------------------
 * sfb_follower_pack_weights re-lays the decoder weights out ONCE PER WEIGHT VERSION into a caller-owned device
 * blob ("shadow copies refreshed when param._version changes"):
 *   - every projection matrix as bf16 (hi, lo) pairs in tcgen05's shared-memory operand layout, one contiguous
 *     32 KB block per (128-row tile, 64-wide K block), LSTM gates interleaved so a tile owns whole cells;
 *   - linear_in_h/linear_in_v of the visual attention folded into M_q = W_v^T W_h (model.py:316-320) and
 *     EltwiseProdScoring folded into M_g = W_a^T diag(w_o) W_h' plus one constant row (model.py:348-351).
 * sfb_follower_step_packed_fwd then runs the step as: [q projection ->] attention gather (+ packs the gate GEMM's
 * activations) -> gate GEMM + LSTM cell -> [t | W_out_h h | next q] projection -> text attention -> h~ projection
 * -> g projection -> action logits [+ rollout tail] (7-8 launches, all projections on tcgen05 fed by bulk async
 * copies); with ctx_k/ctx_o: ... -> gate GEMM + LSTM cell -> W_out_h h || text attention (forms h~) -> [g | next q]
 * -> action logits + tail (6 launches).  Same arguments, outputs and workspace as sfb_follower_step_fwd; `wl` is still needed for the LSTM
 * biases.  Requires H % 128 == 0 and E, F % 8 == 0 (sfb_follower_packed_bytes returns 0 otherwise -> use
 * sfb_follower_step_fwd).  Three optional extras, all NULL-able:
 *   q_in   [B,F]: the visual query W_v^T (W_h h0 + b_h) of THIS step if the caller already has it (the q_next of the
 *                 call that produced h0) — skips the q projection, the first kernel of the dependency chain;
 *   q_next [B,F]: receives the query for the NEXT step (computed from h1 in the same launch as the text-attention
 *                 projection; in train mode, where that launch sees the dropped h1, by one extra launch);
 *   tail        : arguments of sfb_follower_step_tail — the rollout tail (follower.py:476-505) runs fused behind the
 *                 logits in the last kernel; `logit` is then masked in place exactly as by the separate call;
 *   act         : gather source of the action candidates (then `all_u_t` may be NULL);
 *   ctx_k, ctx_o: per-episode projections of ctx from sfb_follower_project_ctx (both or neither). */
/* Where a step's action-candidate embeddings come from (replaces Seq2SeqAgent._action_variable, follower.py:300-320
 * + _build_action_embedding, env.py:60-75).  dense: all_u_t [B,A,E] on the device.  gather: all_u_t = NULL and
 * candidate a of batch row b is  [ feat_table[vp_idx[b], cand_view[b,a], 0:img_dim] , sin(rh) x n, cos(rh) x n,
 * sin(re) x n, cos(re) x n ]  with n = (E - img_dim)/4 and cand_trig[b,a] = {sin rh, cos rh, sin re, cos re}
 * (computed by the host exactly as env.py:68-74 does); cand_view < 0 (the stop action, padding) = all zeros.
 * The candidate rows are views of the slab the attention gather reads, so nothing but indices and 4 floats per
 * candidate crosses PCIe. */
typedef struct sfb_action_source {
  const float*   all_u_t;     /* [B,A,E] or NULL */
  const float*   feat_table;  /* [n_viewpoints, V, img_dim] */
  const int32_t* vp_idx;      /* [B] */
  const int32_t* cand_view;   /* [B,A] view index 0..V-1, or -1 */
  const float*   cand_trig;   /* [B,A,4] */
  int32_t        img_dim;
} sfb_action_source;

typedef struct sfb_step_tail {
  const float*   is_valid;      /* [B,A] */
  const int32_t* target;        /* [B] or NULL */
  int32_t        feedback;      /* 0 teacher, 1 argmax, 2 sample */
  const float*   sample_u;      /* [B] uniforms (feedback == 2) */
  int32_t*       a_t;           /* [B] */
  float*         u_next;        /* [B,E] or NULL */
  float*         action_score;  /* [B] or NULL */
  float*         ce;            /* [B] or NULL */
} sfb_step_tail;
/* Backward of ONE follower decode step: torch autograd of AttnDecoderLSTM.forward (model.py:377-397) as driven by
 * Seq2SeqAgent.train (follower.py:1001-1020, loss.backward() at :1018), hand-written.
 *   forward inputs  : as for sfb_follower_step_fwd (u_prev, action candidates, visual source, h0, c0, ctx, mask, masks)
 *   saved results   : c1, alpha, alpha_v (forward outputs) and `fwd_workspace`, the workspace buffer the forward call of
 *                     THIS step ran in (it still holds feature, activated gates, dropped h1 and h~; a training rollout
 *                     therefore gives every step its own workspace)
 *   upstream grads  : g_h1, g_c1 [B,H], g_logit [B,A] (any may be NULL = zero)
 *   outputs         : d_h0, d_c0 [B,H], d_ctx [B,L,H] (overwritten), parameter gradients in `grads` (state_dict layouts;
 *                     accumulate != 0: added to what is there, as autograd does across the steps of a rollout).
 * u_prev and the action candidates are treated as constants (follower.py:502 detaches them); linear_in_v.bias gets no
 * gradient (it cancels in the softmax).  Every product runs on this library's kernels (skinny GEMMs, attention /
 * LSTM / scoring backward kernels of backward.cu); workspace: sfb_follower_step_bwd_workspace_bytes. */
typedef struct sfb_follower_grads {
  float *lstm_w_ih, *lstm_w_hh, *lstm_b_ih, *lstm_b_hh;
  float *va_w_h, *va_b_h, *va_w_v;
  float *w_in, *w_out;
  float *sc_w_h, *sc_b_h, *sc_w_a, *sc_b_a, *sc_w_out, *sc_b_out;
} sfb_follower_grads;
size_t  sfb_follower_step_bwd_workspace_bytes(const sfb_dims* dims, int32_t B, int32_t L, int32_t A);
int32_t sfb_follower_step_bwd(const sfb_dims* dims, const sfb_vis_lstm_weights* wl,
                              const sfb_softdot_weights* wt, const sfb_scoring_weights* ws,
                              int32_t B, int32_t L, int32_t A,
                              const float* u_prev, const sfb_action_source* act, const sfb_visual_source* vis,
                              const float* h0, const float* c0, const float* ctx, const uint8_t* ctx_mask,
                              const float* drop_x, const float* drop_h,
                              const float* c1, const float* alpha, const float* alpha_v, const void* fwd_workspace,
                              const float* g_h1, const float* g_c1, const float* g_logit,
                              float* d_h0, float* d_c0, float* d_ctx,
                              const sfb_follower_grads* grads, int32_t accumulate,
                              void* workspace, size_t workspace_bytes, void* stream);

size_t  sfb_follower_packed_bytes(const sfb_dims* dims);
size_t  sfb_follower_carry_bytes(const sfb_dims* dims, int32_t B);
int32_t sfb_follower_pack_weights(const sfb_dims* dims, const sfb_vis_lstm_weights* wl,
                                  const sfb_softdot_weights* wt, const sfb_scoring_weights* ws,
                                  void* packed, size_t packed_bytes, void* stream);
int32_t sfb_follower_step_packed_fwd(const sfb_dims* dims, const sfb_vis_lstm_weights* wl,
                                     const void* packed, size_t packed_bytes,
                                     int32_t B, int32_t L, int32_t A,
                                     const float* u_prev, const float* all_u_t, const sfb_visual_source* vis,
                                     const float* h0, const float* c0, const float* ctx, const uint8_t* ctx_mask,
                                     const float* drop_x, const float* drop_h,
                                     float* h1, float* c1, float* alpha, float* logit, float* alpha_v,
                                     void* carry_in, void* carry_out, const sfb_step_tail* tail,
                                     const sfb_action_source* act, const float* ctx_k, const float* ctx_o,
                                     void* workspace, size_t workspace_bytes, void* stream);

/* The first half of a decode step alone — VisualSoftDotAttention + nn.LSTMCell (model.py:389-393) from carried state — as
 * ONE launch (vis_lstm_fused_kernel): the first half of the one-launch step kernel that sfb_follower_step_packed_fwd
 * enqueues when it is given carry_in and the ctx projections, as a launch of its own.  Used by bench.py to time that half
 * in isolation (roofline_first_half) and usable as a building block (the speaker encoder step has the same shape).
 * B <= 128, F = 2176.  feature may be NULL. */
int32_t sfb_follower_gather_lstm_fwd(const sfb_dims* dims, const sfb_vis_lstm_weights* wl,
                                     const void* packed, size_t packed_bytes, int32_t B, void* carry_in,
                                     const sfb_visual_source* vis, const float* c0, float* h1, float* c1,
                                     float* feature, float* alpha_v,
                                     void* workspace, size_t workspace_bytes, void* stream);

/* Per-EPISODE projections of the encoder context (ctx is constant over the decode steps of a rollout,
 * follower.py:446-473): ctx_k = ctx W_in (so that SoftDotAttention's scores ctx . (W_in h) = ctx_k . h, model.py:129-
 * 132) and ctx_o = ctx W_out_c^T (so that W_out_c (sum alpha ctx) = sum alpha ctx_o, model.py:139-141).  Passing both
 * to sfb_follower_step_packed_fwd removes every projection between the LSTM cell and the text attention from the
 * step's dependency chain: W_out_h h is a small projection that runs CONCURRENTLY with the attention (programmatic
 * dependent launch with a deferred dependency wait), tanh is applied while the next projection loads its operand.
 * ctx, ctx_k, ctx_o: [B, L, H]; workspace: sfb_follower_project_ctx_workspace_bytes() (zero-filled once). */
size_t  sfb_follower_project_ctx_workspace_bytes(const sfb_dims* dims, int32_t B, int32_t L);
int32_t sfb_follower_project_ctx(const sfb_dims* dims, const void* packed, size_t packed_bytes,
                                 int32_t B, int32_t L, const float* ctx,
                                 const int32_t* rows, int32_t n_rows,   /* optional: flat (b*L + l) indices of the
                                                                           un-padded positions; the others are neither
                                                                           read nor written (the attention never
                                                                           fetches masked rows); NULL = all B*L */
                                 float* ctx_k, float* ctx_o,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* Table-driven navigation environment: R2RBatch.step + R2RBatch.observe (env.py:628-641, 763-804)
 * as look-ups.  A world state is discretised to (viewpoint, heading bin[, elevation bin]) = one of S ids; for every id
 * the tables hold what `observe` would return for it: the feature-table row and view index of the slab, the candidate
 * list (view index of each navigable direction, sin/cos of its relative heading / elevation — env.py:60-75 —, how many
 * are valid), the successor state of every candidate, and optionally the teacher action per goal (env.py:742-761).
 * sfb_nav_step: rows that have not ended first move along a_prev (a_prev = NULL on the first call; a_prev[b] = 0 ends
 * the row AFTER it is logged, follower.py:531-533; the action is recorded in actions_log, -1 for ended rows), then the
 * step inputs of every row are written: vp_idx, view_idx, cand_view [B,A], cand_trig [B,A,4], is_valid [B,A] and the
 * teacher target [B] (-1 for ended rows or without a teach table).  One tiny launch, no host involvement. */
typedef struct sfb_nav_tables {
  const int32_t* vp;      /* [S] feature-table row */
  const int32_t* view;    /* [S] view index 0..35 */
  const int32_t* nvalid;  /* [S] number of candidates (incl. stop) */
  const int32_t* cv;      /* [S,A] candidate view index, -1 = stop / padding */
  const float*   trig;    /* [S,A,4] */
  const int32_t* next;    /* [S,A] successor state id */
  const int32_t* teach;   /* [S,G] teacher action or NULL */
  int32_t        S, A, G;
} sfb_nav_tables;
int32_t sfb_nav_step(const sfb_nav_tables* t, int32_t B, int32_t* state, int32_t* ended, const int32_t* goal,
                     const int32_t* a_prev, int32_t* actions_log,
                     int32_t* vp_idx, int32_t* view_idx, int32_t* cand_view, float* cand_trig, float* is_valid,
                     int32_t* target, void* stream);

/* State-factored search bookkeeping on the device — follower.py:886-924 for successor_size = 1 over the
 * table-driven environment, where the world-state key (scan, viewpoint, heading, elevation; follower.py:739,894) is the
 * state id of sfb_nav_tables.  The reference's per-instance dicts become dense arrays over the S states: cache (best open
 * inference state per world state), holding (best finished one), completed; inference states live in a per-instance node
 * pool.  One call per search iteration, after the decode step of the states selected by the previous call:
 *   lp [B,A] = log_softmax of the step's masked logits (808-810); `iter` = iteration index (the step's h / c / alpha are
 *   kept by the caller in slot iter + 1), or -1 to use the device's own count flags[3] — a launch replayed from a CUDA
 *   graph cannot take a new argument;
 *   the successors of beam_node[b] are inserted where they strictly improve their table entry (896-900; a successor is
 *   finished after the stop action or at episode_len, 895), then the best not-yet-expanded entry is selected (903-908): an
 *   open one becomes beam_node[b], a finished one moves to completed (912-916); instances with completion_size
 *   completions stop (889-891, 921).  flags[0] is raised when no instance has a state left to expand (925-926) and makes
 *   further calls no-ops, so the host may look at it every few iterations only.  flags[2]: node pool exhausted.
 * All arrays are caller-owned device memory; scores start at -inf, nodes / flags at 0, beam_node at the root node. */
typedef struct sfb_sf_search_state {
  int32_t* beam_node;                                   /* [B] */
  float* c_score; int32_t* c_node; uint8_t* c_exp;     /* [B,S] cache */
  float* h_score; int32_t* h_node; uint8_t* h_exp;     /* [B,S] holding */
  float* d_score; int32_t* d_node; int32_t* n_done;    /* [B,S], [B,S], [B] completed */
  int32_t* n_nodes;                                     /* [B] nodes in use */
  int32_t *node_parent, *node_state, *node_action, *node_count, *node_slot;   /* [B,max_nodes] */
  float* node_score;                                    /* [B,max_nodes] */
  int32_t* trav;                                        /* [B,max_iter] node selected after iteration t, or -1 */
  int32_t* flags;                                       /* [4] ended, scratch, pool overflow, iterations done */
  int32_t max_nodes, max_iter;
} sfb_sf_search_state;
int32_t sfb_sf_search_update(const sfb_sf_search_state* st, const sfb_nav_tables* nav, int32_t B, int32_t iter,
                             int32_t episode_len, int32_t completion_size, const float* lp, void* stream);

/* Per-step tail of Seq2SeqAgent._rollout_with_loss — follower.py:476-505.
 *   logit [B,A] is masked IN PLACE with -inf where is_valid == 0 (477);
 *   feedback: 0 = teacher (a_t = max(target,0)), 1 = argmax, 2 = sample (inverse CDF of
 *   softmax(logit)*valid with the caller's uniforms sample_u [B]);
 *   target [B] int32 (-1 = ignore) or NULL;
 *   -> a_t [B] int32, u_next [B,E] = all_u_t[b, a_t[b]] (502), action_score [B] = log_softmax(logit)[a_t]
 *      (504), ce [B] = -log_softmax(logit)[target] or 0 where target < 0 (481; caller averages). */
int32_t sfb_follower_step_tail(int32_t B, int32_t A, int32_t E, float* logit, const float* is_valid,
                               const int32_t* target, int32_t feedback, const float* sample_u,
                               const float* all_u_t, int32_t* a_t, float* u_next, float* action_score,
                               float* ce, void* stream);

/* EncoderLSTM.forward — model.py:81-104: embedding lookup, 1-layer (optionally bidirectional) LSTM over a
 * length-sorted padded batch with packed-sequence semantics (rows stop at their own length; ctx is zero
 * beyond it), decoder_init = tanh(encoder2decoder(h_T)); c_T returned raw.
 *   seq [B, maxlen] int32 token ids (maxlen = max(lengths)); lengths [B] int32 (any order is accepted);
 *   drop_embed [B*maxlen, Ew] scaled keep mask or NULL (the reference drops embeddings only without GloVe);
 *   Hd = hidden size per direction, H = ndir*Hd.
 *   -> ctx [B, maxlen, H] (un-dropped; forward half then reverse half), decoder_init [B,H], c_t [B,H]
 *      (bidirectional: cat(reverse, forward), model.py:93-94). */
typedef struct sfb_encoder_weights {
  const float* embedding;   /* [vocab, Ew] */
  const float* w_ih[2];     /* [4Hd, Ew]  lstm.weight_ih_l0, lstm.weight_ih_l0_reverse (or NULL) */
  const float* w_hh[2];     /* [4Hd, Hd] */
  const float* b_ih[2];     /* [4Hd] */
  const float* b_hh[2];     /* [4Hd] */
  const float* e2d_w;       /* [H, H] encoder2decoder.weight */
  const float* e2d_b;       /* [H] */
} sfb_encoder_weights;

size_t sfb_encoder_lstm_workspace_bytes(int32_t ndir, int32_t Hd, int32_t Ew, int32_t B, int32_t maxlen);

/* Same as sfb_encoder_lstm_fwd, with the number of rows of w->embedding stated: without dropout on the embedding
 * (drop_embed == NULL) the input projection is then computed once per VOCABULARY row instead of once per (row, time)
 * position (nn.Embedding + the W_ih half of nn.LSTM, model.py:62-66,85-90).  Same results. */
int32_t sfb_encoder_lstm_fwd_vocab(const sfb_encoder_weights* w, int32_t vocab, int32_t ndir, int32_t Hd, int32_t Ew,
                                   int32_t B, int32_t maxlen, const int32_t* seq, const int32_t* lengths,
                                   const float* drop_embed, float* ctx, float* decoder_init, float* c_t,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* Training (autograd of EncoderLSTM.forward, model.py:81-104, unidirectional): sfb_encoder_lstm_train_fwd is the same
 * forward that additionally keeps a TAPE (activated gates and the (h, c) state around every time step, caller-owned,
 * sfb_encoder_lstm_tape_bytes); sfb_encoder_lstm_bwd back-propagates through time over it, hand-written:
 *   g_ctx [B,maxlen,H] / g_decoder_init [B,H] / g_c_t [B,H]: upstream gradients (NULL = zero);
 *   grads: state_dict layouts (lstm.weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0, encoder2decoder.weight/.bias),
 *   accumulated when accumulate != 0.  The embedding is frozen GloVe in the reference configuration (model.py:58-60)
 *   and receives no gradient here. */
typedef struct sfb_encoder_grads { float *w_ih, *w_hh, *b_ih, *b_hh, *e2d_w, *e2d_b; } sfb_encoder_grads;
size_t  sfb_encoder_lstm_tape_bytes(int32_t Hd, int32_t B, int32_t maxlen);
int32_t sfb_encoder_lstm_train_fwd(const sfb_encoder_weights* w, int32_t Hd, int32_t Ew, int32_t B, int32_t maxlen,
                                   const int32_t* seq, const int32_t* lengths, const float* drop_embed,
                                   float* ctx, float* decoder_init, float* c_t, void* tape, size_t tape_bytes,
                                   void* workspace, size_t workspace_bytes, void* stream);
size_t  sfb_encoder_lstm_bwd_workspace_bytes(int32_t Hd, int32_t Ew, int32_t B, int32_t maxlen);
int32_t sfb_encoder_lstm_bwd(const sfb_encoder_weights* w, int32_t Hd, int32_t Ew, int32_t B, int32_t maxlen,
                             const int32_t* seq, const int32_t* lengths, const float* drop_embed, const void* tape,
                             const float* decoder_init, const float* g_ctx, const float* g_decoder_init, const float* g_c_t,
                             const sfb_encoder_grads* grads, int32_t accumulate,
                             void* workspace, size_t workspace_bytes, void* stream);

int32_t sfb_encoder_lstm_fwd(const sfb_encoder_weights* w, int32_t ndir, int32_t Hd, int32_t Ew, int32_t B,
                             int32_t maxlen, const int32_t* seq, const int32_t* lengths, const float* drop_embed,
                             float* ctx, float* decoder_init, float* c_t,
                             void* workspace, size_t workspace_bytes, void* stream);

/* SpeakerEncoderLSTM._forward_one_step — model.py:429-435 (visual attention + LSTMCell). */
int32_t sfb_speaker_encoder_step_fwd(const sfb_dims* dims, const sfb_vis_lstm_weights* w, int32_t B,
                                     const float* action_embedding, const sfb_visual_source* vis,
                                     const float* h0, const float* c0, const float* drop_x,
                                     float* h1, float* c1,
                                     void* workspace, size_t workspace_bytes, void* stream);

/* SpeakerDecoderLSTM.forward (default branch) — model.py:497-503,515-519.
 *   prev_word [B] int32, embedding [vocab, Ew], LSTMCell Ew->H, SoftDotAttention over ctx [B,T,H]
 *   with mask [B,T], vocabulary projection w_voc [vocab, H] + b_voc -> logit [B, vocab]. */
typedef struct sfb_speaker_decoder_weights {
  const float* embedding;   /* [vocab, Ew] embedding.weight */
  const float* lstm_w_ih;   /* [4H, Ew] */
  const float* lstm_w_hh;   /* [4H, H] */
  const float* lstm_b_ih;   /* [4H] */
  const float* lstm_b_hh;   /* [4H] */
  sfb_softdot_weights attn; /* attention_layer.* */
  const float* w_voc;       /* [vocab, H] decoder2action.weight */
  const float* b_voc;       /* [vocab] */
} sfb_speaker_decoder_weights;

size_t sfb_speaker_decoder_step_workspace_bytes(int32_t H, int32_t Ew, int32_t B, int32_t T);

int32_t sfb_speaker_decoder_step_fwd(const sfb_speaker_decoder_weights* w, int32_t H, int32_t Ew, int32_t vocab,
                                     int32_t B, int32_t T, const int32_t* prev_word,
                                     const float* h0, const float* c0, const float* ctx, const uint8_t* ctx_mask,
                                     const float* drop_e, const float* drop_h,
                                     float* h1, float* c1, float* alpha, float* logit,
                                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- packed-weight fast paths of the two speaker steps (same idea as sfb_follower_pack_weights) ------------------
 * sfb_vis_lstm_pack_weights: folded visual query M_q = W_v^T W_h + gate-interleaved [W_ih | W_hh] as tcgen05 operand
 * tiles; sfb_speaker_encoder_step_packed_fwd = SpeakerEncoderLSTM._forward_one_step (model.py:429-435) in 3 launches
 * (q projection, attention gather that also packs the gate GEMM's activations, gate GEMM + LSTM cell).
 * sfb_speaker_decoder_pack_weights / _step_packed_fwd = SpeakerDecoderLSTM.forward (model.py:497-503,515-519) in
 * 5 launches, every projection on tcgen05 (the embedding row of the previous word is gathered and split to bf16
 * hi/lo while the gate GEMM loads its operand).  Same arguments, outputs and workspaces as the in-place entry points;
 * *_packed_bytes return 0 for dimensions the packed path does not cover (H % 128, E/F % 8, Ew % 4, vocab <= 4096). */
size_t  sfb_vis_lstm_packed_bytes(const sfb_dims* dims);
int32_t sfb_vis_lstm_pack_weights(const sfb_dims* dims, const sfb_vis_lstm_weights* w,
                                  void* packed, size_t packed_bytes, void* stream);
int32_t sfb_speaker_encoder_step_packed_fwd(const sfb_dims* dims, const sfb_vis_lstm_weights* w,
                                            const void* packed, size_t packed_bytes, int32_t B,
                                            const float* action_embedding, const sfb_visual_source* vis,
                                            const float* h0, const float* c0, const float* drop_x,
                                            float* h1, float* c1,
                                            void* workspace, size_t workspace_bytes, void* stream);
size_t  sfb_speaker_decoder_packed_bytes(int32_t H, int32_t Ew, int32_t vocab);
int32_t sfb_speaker_decoder_pack_weights(const sfb_speaker_decoder_weights* w, int32_t H, int32_t Ew, int32_t vocab,
                                         void* packed, size_t packed_bytes, void* stream);
int32_t sfb_speaker_decoder_step_packed_fwd(const sfb_speaker_decoder_weights* w, const void* packed, size_t packed_bytes,
                                            int32_t H, int32_t Ew, int32_t vocab, int32_t B, int32_t T,
                                            const int32_t* prev_word,
                                            const float* h0, const float* c0, const float* ctx, const uint8_t* ctx_mask,
                                            const float* drop_e, const float* drop_h,
                                            float* h1, float* c1, float* alpha, float* logit,
                                            void* workspace, size_t workspace_bytes, void* stream);

/* Number of kernels the last successful call on this thread enqueued (bench.py's gpu_launches). */
int32_t sfb_last_launch_count(void);

/* ---- speaker modules under autograd (train_speaker.py:67-118 -> speaker.py:376-395 loss.backward()).
 * Hand-written backward of SpeakerEncoderLSTM._forward_one_step (model.py:429-435) and of SpeakerDecoderLSTM.forward
 * (model.py:487-519, default branch), on the same kernels as sfb_follower_step_bwd.  fwd_workspace: the workspace the
 * step's FORWARD call ran in (sfb_speaker_encoder_step_fwd / _packed_fwd, sfb_speaker_decoder_step_fwd / _packed_fwd):
 * it still holds the attention output, the attention weights, the activated gates, the dropped h_1 and h~.
 * Gradients w.r.t. parameters are written to (accumulate == 0) or added to (accumulate != 0) the state_dict-shaped
 * buffers of `grads` (NULL = skip); gradients w.r.t. the recurrent state (and the decoder's ctx) are returned.  The
 * word embedding is frozen GloVe in the reference configuration (model.py:469-472) and receives no gradient; the
 * action embeddings and image features are data. */
size_t  sfb_speaker_encoder_step_bwd_workspace_bytes(const sfb_dims* dims, int32_t B);
int32_t sfb_speaker_encoder_step_bwd(const sfb_dims* dims, const sfb_vis_lstm_weights* w, int32_t B,
                                     const float* action_embedding, const sfb_visual_source* vis, const float* h0,
                                     const float* c0, const float* drop_x, const float* c1, const void* fwd_workspace,
                                     const float* g_h1, const float* g_c1, float* d_h0, float* d_c0,
                                     const sfb_follower_grads* grads /* lstm_* and va_* fields */, int32_t accumulate,
                                     void* workspace, size_t workspace_bytes, void* stream);
typedef struct sfb_speaker_decoder_grads {
  float *lstm_w_ih, *lstm_w_hh, *lstm_b_ih, *lstm_b_hh;
  float *w_in, *w_out;      /* attention_layer.linear_in / linear_out */
  float *w_voc, *b_voc;     /* decoder2action */
} sfb_speaker_decoder_grads;
size_t  sfb_speaker_decoder_step_bwd_workspace_bytes(int32_t H, int32_t Ew, int32_t vocab, int32_t B);
int32_t sfb_speaker_decoder_step_bwd(const sfb_speaker_decoder_weights* w, int32_t H, int32_t Ew, int32_t vocab,
                                     int32_t B, int32_t T, const int32_t* prev_word, const float* h0, const float* c0,
                                     const float* ctx, const uint8_t* ctx_mask, const float* drop_e, const float* drop_h,
                                     const float* c1, const float* alpha, const void* fwd_workspace,
                                     const float* g_h1, const float* g_c1, const float* g_logit,
                                     float* d_h0, float* d_c0, float* d_ctx, const sfb_speaker_decoder_grads* grads,
                                     int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SF_B200_H_ */

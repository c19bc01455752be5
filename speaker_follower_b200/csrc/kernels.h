// kernels.h — internal kernel parameter blocks and launchers (not part of the C ABI).
#pragma once
#include "../../include/sf_b200.h"
#include "common.cuh"

namespace sfb {

// ---------------------------------------------------------------- pack.cu
struct PackSeg {
  const float* x; int ldx; int k;
  const float* xs; int ldxs;             // optional elementwise scale, indexed by the logical row
  const int32_t* xrow;                   // optional row indirection
};
struct PackParams {
  PackSeg seg[3];
  int nseg;
  int ntile;                             // row tiles (weights: ceil(N/128) or H/32; activations: batch tiles)
  int R;                                 // packed rows per tile (128 for weights, NB for activations)
  int rows_per_tile, rows_valid;         // plain mapping: source row = tile*rows_per_tile + r, valid below rows_valid
  int lstm_H;                            // > 0: gate-interleaved LSTM weights, source row = (r/32)*H + tile*32 + r%32
  unsigned char* out;
  int nkb;                               // filled by the launcher
  unsigned long long* trace;
};
int32_t pack_prepare(PackParams& p);   // validates and fills nkb
int32_t launch_pack_rows(const PackParams& p, cudaStream_t stream);
int32_t launch_transpose(const float* src, int lds, int R, int Cc, float* dst, int ldd, cudaStream_t stream);
int32_t launch_fold(const float* A, int lda, const float* s, const float* Bm, int ldb, const float* bv, int D, int NA,
                    int NJ, float* out, int ldo, float* obias, const float* obias_add, cudaStream_t stream);

// ---------------------------------------------------------------- attention.cu
struct AttnParams {
  const float* q;   int ldq;              // [B, D] query
  // row r of batch element b = concat(segA[...], segB[...]); lenA + lenB == D
  const float* segA; long long strideA_b; int strideA_r; int lenA;
  const float* segB; long long strideB_b; int strideB_r; int lenB;
  // optional separate KEY rows (same length D, dense [B,R,D] addressing): scores = key . q, output = sum alpha * row
  const float* keyA; long long strideK_b; int strideK_r;
  const int32_t* idxA;                    // optional: batch element b reads block idxA[b] of segA (gather)
  const int32_t* idxB;
  const uint8_t* mask; int ldmask;        // [B, R], 1 = masked; may be NULL
  int R, D;
  float* out;   int ldo;                  // [B, D]
  float* alpha; int ldalpha;              // [B, R] or NULL
  // optional second output: `out` (times an elementwise scale = dropout keep mask) written as bf16 hi/lo into the
  // packed activation operand of the following gate GEMM (layout: pack.cu), K blocks pk_kb0.. of pk_nkb
  unsigned char* pk_out; int pk_kb0, pk_nkb, pk_NB, pk_rows_per_z;
  const float* pk_scale; int pk_ldscale;
  int has_side; PackParams side;          // optional side job after the dependency wait: pack other step operands
  // 1: the query was produced TWO kernels back and the predecessor (which triggers its dependents only after its own
  // dependency wait) writes nothing this kernel reads -> run concurrently with it; the dependency wait moves to the end
  // of the kernel so that this grid's completion still implies the predecessor's
  int defer_wait;
  const float* post_add; int ld_post; int post_tanh;   // out = tanh?(attention output + post_add[b, :]) (read after the wait)
  int idx_dependent;                      // 1: idxA / idxB are written by the preceding kernel: read them after the dependency wait
  int no_hint;                            // bring-up: plain L2 policy instead of evict-first
  unsigned long long* cta_trace;          // bring-up: per-CTA {entry, first row landed, stream done, exit, smid}
  unsigned long long* trace;              // bring-up: 3 timestamps of block 0, or NULL
  // filled by the launcher
  int rows_per_cta, stages;
  float* part;                            // [B][SPLIT][D+4] partial (max, sum, weighted sum) records
  unsigned int* ticket;                   // [B] self-resetting arrival counters (must start zeroed)
};
struct AttnPlan {
  int split, rows_per_cta, stages;
  size_t ticket_bytes, bytes;             // workspace: tickets + partial records
};
AttnPlan attention_plan(int B, int R, int D, int num_sms, bool kv = false);
int32_t launch_soft_dot_attention(AttnParams p, int B, void* ws, size_t ws_bytes, cudaStream_t stream);

// ---------------------------------------------------------------- gemm_simt.cu
// out[M,N] = act( sum_s (x_s .* xs_s) · w_s^T + bias0 + bias1 )   with M <= a few hundred ("skinny").
struct GemmSeg {
  const float* x;  int ldx;               // [M, k] activations (row stride ldx)
  const int32_t* xrow;                    // optional row indirection: row m reads x[xrow[m]] (embedding lookup)
  const float* xs; int ldxs;              // optional elementwise scale (dropout keep mask); ldxs may be 0
  const float* w;  int ldw;               // w_kn == 0: [N, k] row-major (nn.Linear); w_kn == 1: [k, N] row-major
  int k;
  int w_kn;
  // gemm_pk only: the activation actually multiplied is f(x + xadd) with f = tanh when xtanh (fuses the
  // h~ = tanh(W_out_c wc + W_out_h h) pointwise step of SoftDotAttention into the next projection's operand load)
  const float* xadd; int ldxadd; int xtanh;
};
// Fused LSTM cell epilogue (nn.LSTMCell pointwise part, model.py:393): active when H > 0.  The GEMM then
// computes the gate pre-activations W_ih x + W_hh h with gate-interleaved 32-column tiles (4 gates x 8 units).
struct LstmEpilogue {
  int H;                                  // 0 = plain epilogue
  const float* b_ih; const float* b_hh;   // [4H]
  const float* c0;                        // [M,H]
  const float* drop_h;                    // [M,H] scaled keep mask or NULL
  float* h1; float* c1;                   // [M,H]
  float* h1_drop;                         // [M,H] h1 .* drop_h (== h1 when drop_h is NULL); may be NULL
  float* gates_act;                       // [M,4H] activated gates (i,f,g,o) kept for backward; may be NULL
  // sequence mode (EncoderLSTM, model.py:89-90): precomputed input projection + packed-sequence masking
  const float* addend; long long ld_addend;   // [M,4H] rows at stride ld_addend (x_t W_ih^T), or NULL
  const int32_t* addend_rows;                 // optional: row m reads addend row addend_rows[m] (a per-token table)
  const float* h0;                            // previous hidden state, carried through when t >= lengths[m]
  const int32_t* lengths; int t;              // rows with t >= lengths[m] keep (h0, c0) and emit zeros
  float* seq_out; long long ld_seq_out;       // h1 (or 0 when inactive) written at seq_out[m*ld_seq_out + j]
  // optional: h1 (un-dropped) also emitted as bf16 (hi, lo) into a packed activation operand (layout: pack.cu) at
  // K blocks hpk_kb0.. of hpk_nkb — the h_0 blocks of the NEXT step's gate GEMM
  unsigned char* hpk; int hpk_kb0, hpk_nkb, hpk_NB, hpk_rows_per_z;
  // optional: h1_drop as packed operand [H/64 blocks][2*hpk_NB*128 B] (one batch tile) for the fused text-side kernel
  unsigned char* hdpk;
};
struct GemmParams {
  GemmSeg seg[3];
  int nseg;
  int M, N;
  int splitk;                             // 1, 2, 4 or 8: K split over the CTAs of a cluster, reduced through DSMEM
  float* out; int ldo;                    // plain epilogue: out = act(acc + bias0 + bias1)
  const float* bias0; const float* bias1; // [N] or NULL
  const float* oscale;                    // [N] or NULL: out = act(acc + biases) * oscale[n]
  const float* padd; int ld_padd;         // [M, >=N] or NULL: added before the activation (pre-computed partial sum)
  // optional split output: columns n >= n_split (a multiple of 128) go to out2[m*ldo2 + n - n_split] + bias2, no act
  int n_split; float* out2; int ldo2; const float* bias2;
  const int32_t* out_rows;                // optional: GEMM row m is stored at output row out_rows[m] (compacted batches)
  int n1_valid;                           // with out2: columns n1_valid <= n < n_split are padding and not stored (0 = n_split)
  int act;                                // 0 none, 1 tanh
  int exact;                              // 1: force the exact-fp32 FFMA path (default: 3xTF32 mma.sync for M > 32)
  LstmEpilogue lstm;
  unsigned long long* trace;              // bring-up: 3 timestamps of block 0, or NULL
};
int32_t launch_gemm(const GemmParams& p, cudaStream_t stream);
int gemm_pick_splitk(int M, int N, int ktotal, int num_sms);

// ---------------------------------------------------------------- gemm_tc.cu (tcgen05 / TMEM, LSTM epilogue only)
struct TcPlan {
  int tiles, nz, rows_per_z, NB, S;
  size_t sem_bytes, bytes;               // workspace: self-resetting semaphores (must start zeroed) + partial tiles
};
TcPlan gemm_tc_plan(int M, int H, int ktotal, int nseg, int num_sms);
bool gemm_tc_supported(const GemmParams& p);
int32_t launch_gemm_tc(const GemmParams& p, cudaStream_t stream, void* ws, size_t ws_bytes);
void gemm_tc_set_debug(int flags);
int gemm_tc_read_timestamps(long long* out, int n);
extern int g_disable_tc;
extern int g_disable_fused;

// ---------------------------------------------------------------- gemm_pk.cu (tcgen05 from pre-packed weights)
struct PkParams {
  GemmParams g;                          // epilogue (plain or LSTM), M, N; seg[] = fp32 activations when b_pk == NULL
  const unsigned char* a_pk;             // packed weights [tiles][nkb][32 KB]
  const unsigned char* b_pk;             // packed activations [nz][nkb][2*NB*128 B] or NULL
  int nkb;                               // 64-wide K blocks (all segments)
  int has_side; PackParams side;         // optional side job for the otherwise idle warps: pack another operand
  int wide;                              // 1: packed activations use 256-row batch tiles when M > 2048 (UMMA N = 256)
  int late_trigger;                      // 1: griddepcontrol.launch_dependents only after the dependency wait (see AttnParams::defer_wait)
  int alt_tile0; GemmSeg alt_seg;        // tiles >= alt_tile0 (> 0) read their fp32 activations from alt_seg instead of g.seg[0]
  // filled by the launcher
  int NB, rows_per_z, nstages;
  float* partial; unsigned int* sem;
};
struct PkPlan {
  int tiles, nz, rows_per_z, NB, S;
  size_t sem_bytes, bytes;               // workspace: self-resetting semaphores (must start zeroed) + partial tiles
};
PkPlan gemm_pk_plan(int M, int N_rows, int nkb, bool b_packed, int num_sms, bool wide = false);
int pk_num_kblocks(const int* seg_k, int nseg);
size_t pk_weight_bytes(int N_rows, int nkb);
size_t pk_act_bytes(int M, int nkb, bool wide = false);
int32_t launch_gemm_pk(const PkParams& q, cudaStream_t stream, void* ws, size_t ws_bytes);
int pk_max_active_clusters(int cluster, int smem);

// ---------------------------------------------------------------- step_fused.cu (attention gather + LSTM cell, one launch)
struct FusedVisLstmParams {
  // gather role: batch element b = one CTA; row r of b = concat(segA[idxA[b]][r], segB[idxB[b]][r]), rows contiguous
  const float* q; int ldq;                // [B, D] visual query
  const float* segA; long long strideA_b; int lenA;
  const float* segB; long long strideB_b; int lenB;
  const int32_t* idxA; const int32_t* idxB;
  int R, D;
  float* feat; int ldfeat;                // [B, D] attention output (fp32)
  float* alpha; int ldalpha;              // [B, R] or NULL
  const float* pk_scale; int pk_ldscale;  // dropout keep-scale of the feature part of the gate operand, or NULL
  // GEMM role: packed weights [tiles][nkb][32 KB], packed activations [nkb][2*NB*128 B] whose blocks
  // [post_kb0, post_kb1) (the attention output) are written by the gather role of this launch
  const unsigned char* a_pk; unsigned char* b_pk; int nkb, post_kb0, post_kb1;
  int feat_kb0;                           // first K block of the attention output inside b_pk
  GemmParams g;                           // LSTM epilogue (g.lstm), M = B
  int B;
  int idx_dependent;                      // 1: idxA / idxB are written by the preceding kernel -> read them after the dependency wait
  int dbg;                                // bring-up: 1 = gather without arithmetic, 2 = dot products only
  int pre_weight_free;                    // share of the pre K blocks a CTA without a gather role takes, relative to 1 for a gather CTA
  // filled by the launcher
  int NB, nch, chunk_rows;
  float* partial; unsigned int* sem; unsigned int* sync;
};
struct FusedPlan {
  bool ok;
  int tiles, S, NB, nch, chunk_rows, gstages;
  size_t smem, sem_bytes, bytes;
  size_t data_bytes;   // leading part of the dynamic shared memory that holds only staged data (free once the MMAs have retired)
};
FusedPlan vis_lstm_fused_plan(int B, int H, int nkb, int R, int D, int lenA, int lenB, int num_sms);
int32_t launch_vis_lstm_fused(const FusedVisLstmParams& q, cudaStream_t stream, void* ws, size_t ws_bytes);

// rollout tail of one batch row (pointwise.cu / tail.cuh)
struct TailParams {
  float* logit; const float* is_valid; const int32_t* target; int feedback; const float* sample_u;
  const float* all_u_t; int32_t* a_t; float* u_next; float* action_score; float* ce;
  int B, A, E;
  unsigned long long* trace;
  // optional: the chosen candidate row also as bf16 (hi, lo) in the packed activation operand (K blocks 0..) of the
  // NEXT step's gate GEMM (layout: pack.cu, one batch tile of upk_NB rows)
  unsigned char* upk; int upk_NB;
};

// ---------------------------------------------------------------- step_fused_b.cu (text attention + projections + scoring + tail, one launch)
struct FusedTextScoreParams {
  int B, L, A, H, E;
  // text attention over the per-episode key / value projections of ctx (sfb_follower_project_ctx)
  const float* h1d; int ldh;                 // [B,H] query: h_1 after dropout
  const float* ctx_k; const float* ctx_o;    // [B,L,H]
  const uint8_t* mask; int ldmask;           // [B,L] 1 = masked, or NULL
  float* alpha; int ldalpha;                 // [B,L] or NULL
  float* h_tilde;                            // [B,H] fp32 (kept for the API / backward) or NULL
  unsigned char* htpk;                       // h~ as packed bf16 (hi, lo) operand [H/64][2*NB*128 B]
  // projections on tcgen05 from packed operands, K = H
  const unsigned char* a_hh; int hh_tiles;   // W_out[:, H:2H] rows
  const unsigned char* hdpk;                 // packed h1d
  float* hh; int ldhh;                       // [B,H] scratch
  const unsigned char* a_q; int q_tiles;     // M_q rows; q_tiles = 0: no next query
  const unsigned char* hpk;                  // packed (un-dropped) h_1
  float* q_next; int ldq; const float* b_q; int q_cols;
  const unsigned char* a_g; int g_tiles;     // M_g rows + constant row
  float* g; int ldg; const float* b_g; int g_cols;
  // action candidates (dense or gathered, see ScoringParams) + logits + rollout tail
  const float* all_u_t;
  const float* cand_table; const int32_t* vp_idx; const int32_t* cand_view; const float* cand_trig; int img_dim, cand_V;
  float* logit;
  int has_tail; TailParams tail;
  // filled by the launcher
  int NB, P, nch, nkb;
  unsigned int* sync;
  unsigned long long* trace;
};
struct FusedTextPlan {
  bool ok;
  int NB, P, grid, nch;
  size_t smem, sync_bytes;
};
FusedTextPlan text_score_fused_plan(int B, int L, int A, int H, int E, int F, bool with_q, int num_sms);
int32_t launch_text_score_fused(const FusedTextScoreParams& q, cudaStream_t stream, void* sync_ws, size_t sync_bytes);

// ---------------------------------------------------------------- encoder_persist.cu
// the recurrent part of EncoderLSTM for all time steps in one launch (W_hh resident in shared memory)
struct EncPersistParams {
  const unsigned char* whh_pk[2];     // per direction: gate-interleaved packed W_hh (pack.cu, lstm_H = Hd)
  const float* xproj[2];              // per direction: hoisted input projection [B * maxlen][4 Hd], or with `seq` the
                                      // per-token table [vocab][4 Hd] = Emb W_ih^T
  const int32_t* seq;                 // [B][maxlen] word ids selecting the rows of xproj, or NULL (rows by position)
  const float* b_ih[2]; const float* b_hh[2];
  const int32_t* lengths;
  float* ctx; long long ld_ctx; int H;   // [B][maxlen][H = ndir * Hd]
  float* h_fin; float* c_fin;         // [ndir][B][Hd] final states
  float* tape_h; float* tape_c; float* tape_g;   // optional training tape: h, c [maxlen + 1][B][Hd], gate activations [maxlen][B][4 Hd]
  unsigned char* hpk;                 // [groups][2 parities][Hd / 64][2 * (N / 8) * 1024] packed h between steps
  unsigned long long* bar;            // [groups] step counters (+ status word)
  int ndir, Hd, B, maxlen;
  // filled by the launcher
  int NG, Bg, N;
  unsigned int* status;
  unsigned long long* trace;
};
struct EncPersistPlan {
  bool ok;
  int NG, Bg, N, grid;
  size_t smem, hpk_bytes, bar_bytes;
};
EncPersistPlan encoder_persist_plan(int ndir, int Hd, int B, int num_sms);
int32_t launch_encoder_persist(const EncPersistParams& q, cudaStream_t stream);

// both halves of the step as one launch (step_fused_b.cu: step_kernel)
struct FusedStepParams {
  FusedVisLstmParams a;
  FusedTextScoreParams b;
  unsigned int* phase;      // arrivals: one per CTA when its share of the first half is visible device-wide
  int top_off, tiles, S;
  int prefetch;             // bring-up option "merged_prefetch"
};
struct FusedStepPlan {
  bool ok;
  FusedPlan a;
  FusedTextPlan b;          // P / nch for the merged geometry
  size_t smem, top_off;
};
FusedStepPlan step_fused_plan(int B, int L, int A, int H, int E, int F, bool with_q, int nkb, int R, int D, int lenA, int lenB, int num_sms);
// ws / ws_bytes: the first half's workspace; sync_ws: the second half's counter words (word 8 = the phase counter)
extern int g_merged_prefetch;
int32_t launch_step_fused(const FusedVisLstmParams& qa, const FusedTextScoreParams& qb, cudaStream_t stream, void* ws, size_t ws_bytes,
                          void* sync_ws, size_t sync_bytes);

// ---------------------------------------------------------------- pointwise.cu
// logit[b,a] = all_u_t[b,a,:] . g[b,:] + sum_d b_a[d] w_out[d] tp[b,d] + b_out   (EltwiseProdScoring rewritten)
struct ScoringParams {
  const float* all_u_t;                   // [B,A,E]
  const float* g;                         // [B,E]
  const float* tp;                        // [B,D]  linear_in_h(h_tilde) (.) w_out; NULL: constant = g[b*ldg + E] (folded weights)
  const float* b_a; const float* b_out;
  int ldg;                                // row stride of g
  int has_tail; TailParams tail;          // optional fused rollout tail (follower.py:476-505) on the fresh logits
  // optional gather source replacing all_u_t (env.py:60-75): candidate a of row b = view cand_view[b,a] of the
  // viewpoint's slab in the feature table + the 4 trig values of its relative heading / elevation; view < 0 = zeros
  const float* cand_table; const int32_t* vp_idx; const int32_t* cand_view; const float* cand_trig; int img_dim, cand_V;
  float* logit;                           // [B,A]
  int B, A, E, D;
  unsigned long long* trace;
};
int32_t launch_action_scoring(const ScoringParams& p, cudaStream_t stream);

int32_t launch_follower_tail(const TailParams& p, cudaStream_t stream);

// table-driven navigation environment: S discretised world states, A candidate slots per state, G goals
struct NavStepParams {
  const int32_t* vp; const int32_t* view; const int32_t* nvalid;   // [S]
  const int32_t* cv; const float* trig; const int32_t* next;       // [S,A], [S,A,4], [S,A]
  const int32_t* teach; const int32_t* goal;                       // [S,G] or NULL, [B]
  int B, A, G;
  int32_t* state; int32_t* ended;                                  // [B] in/out
  const int32_t* a_prev; int32_t* actions_log;                     // [B] or NULL
  int32_t* vp_idx; int32_t* view_idx; int32_t* cand_view; float* cand_trig; float* is_valid; int32_t* target;
};
int32_t launch_nav_step(const NavStepParams& p, cudaStream_t stream);

// state-factored search bookkeeping on the device (pointwise.cu, follower.py:886-924, successor_size = 1)
struct SfSearchParams {
  int B, A, S, M, max_iter, episode_len, completion_size, iter;
  const float* lp;                       // [B,A] log-softmax of this iteration's masked logits
  const int32_t* nav_next; const int32_t* nav_nvalid;
  int32_t* beam_node;                    // [B] in: node expanded this iteration (-1: none); out: node to expand next
  float* c_score; int32_t* c_node; uint8_t* c_exp;    // [B,S] cache
  float* h_score; int32_t* h_node; uint8_t* h_exp;    // [B,S] holding
  float* d_score; int32_t* d_node; int32_t* n_done;   // [B,S], [B] completed
  int32_t* n_nodes;                      // [B]
  int32_t *node_parent, *node_state, *node_action, *node_count, *node_slot; float* node_score;   // [B,M]
  int32_t* trav;                         // [B,max_iter] node selected for expansion after iteration t, or -1
  int32_t* flags;                        // [0] ended, [1] scratch, [2] node pool overflow, [3] iterations done
};
int32_t launch_sf_search_update(const SfSearchParams& p, cudaStream_t stream);

// ---------------------------------------------------------------- backward.cu
struct ScoreBwdParams {
  const float* dlogit; int B, A, E;
  const float* all_u_t;
  const float* cand_table; const int32_t* vp_idx; const int32_t* cand_view; const float* cand_trig; int img_dim, cand_V;
  float* dg; float* dsum;                 // [B,E], [B]
};
int32_t launch_score_bwd(const ScoreBwdParams& p, cudaStream_t st);
int32_t launch_tanh_bwd(const float* dy, const float* y, float* dz, int n, cudaStream_t st);
int32_t launch_scoring_mid(const float* th, const float* v, const float* w_o, const float* b_a, const float* dsum, float* r,
                           float* dth, float* prod, int B, int D, cudaStream_t st);
struct LstmBwdParams {
  int B, H;
  const float* gates_act; const float* c0; const float* c1;
  const float* g_h1; const float* g_c1; const float* dh1d; const float* drop_h;   // any of the three gradients may be NULL
  float* dgates; float* dc0;
};
int32_t launch_lstm_cell_bwd(const LstmBwdParams& p, cudaStream_t st);
// EncoderLSTM BPTT cell step: rows with t >= lengths[row] pass (dh, dc) through untouched (packed sequence)
struct LstmSeqBwdParams {
  int B, H, t; const int32_t* lengths;
  const float* gates_act; const float* c_prev; const float* c_cur;
  const float* dh_in; const float* dc_in;
  const float* g_out; long long ld_g_out;   // gradient of the emitted h_t (ctx[:, t, :]) or NULL
  float* dgates; float* dc_prev; float* dh_pass;
};
int32_t launch_lstm_seq_bwd(const LstmSeqBwdParams& p, cudaStream_t st);
int32_t launch_gather_embed(const float* emb, int Ew, const int32_t* seq, const float* drop, float* out, int B, int maxlen, cudaStream_t st);
int32_t launch_assemble_x(const float* u, const float* f, const float* drop, float* x, int B, int E, int F, cudaStream_t st);
struct AttnBwdParams {
  const float* segA; long long strideA_b; int strideA_r, lenA;     // rows as in AttnParams (dense or gathered)
  const float* segB; long long strideB_b; int strideB_r, lenB;
  const int32_t* idxA; const int32_t* idxB;
  const uint8_t* mask; int ldmask;
  int R, D;
  const float* alpha; int ldalpha;        // forward softmax weights [B,R]
  const float* dout; int lddout;          // gradient of the weighted sum [B,D]
  const float* dout_scale; int ldscale;   // optional elementwise factor of dout (dropout keep mask)
  const float* qv; int ldq;               // query of the scores (only read when drows != NULL)
  float* dq; int lddq;                    // [B,D] gradient w.r.t. the query
  float* drows;                           // [B,R,D] gradient w.r.t. the rows, or NULL
  float* wsum; int ldwsum;                // [B,D] the forward weighted sum (recomputed), or NULL
  int stage_rows;                         // filled by the launcher
};
int32_t launch_attn_bwd(AttnBwdParams p, int B, cudaStream_t st);
struct OuterParams {
  const float* Y; int ldy; const float* X; int ldx;
  int B, N, K;
  float* out; int ldo; int accumulate;
};
int32_t launch_outer_accum(const OuterParams& p, cudaStream_t st);
int32_t launch_colsum(const float* Y, int ldy, int B, int N, float* out, int accumulate, cudaStream_t st);

int device_num_sms();
unsigned long long* next_trace_slot();
unsigned long long* cta_trace_buffer();  // NULL unless sfb_set_option("cta_trace", 1)
extern int g_attn_force_cl, g_attn_force_stages, g_attn_no_hint, g_attn_ring_kb, g_attn_rb;   // NULL unless sfb_set_option("trace", 1)

}  // namespace sfb

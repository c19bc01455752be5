// kernels.h — internal kernel parameter blocks and launchers (not part of the C ABI).
#pragma once
#include "../../include/sf_b200.h"
#include "common.cuh"

namespace sfb {

// ---------------------------------------------------------------- attention.cu
struct AttnParams {
  const float* q;   int ldq;              // [B, D] query
  // row r of batch element b = concat(segA[...], segB[...]); lenA + lenB == D
  const float* segA; long long strideA_b; int strideA_r; int lenA;
  const float* segB; long long strideB_b; int strideB_r; int lenB;
  const int32_t* idxA;                    // optional: batch element b reads block idxA[b] of segA (gather)
  const int32_t* idxB;
  const uint8_t* mask; int ldmask;        // [B, R], 1 = masked; may be NULL
  int R, D;
  float* out;   int ldo;                  // [B, D]
  float* alpha; int ldalpha;              // [B, R] or NULL
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
AttnPlan attention_plan(int B, int R, int D, int num_sms);
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
  const float* h0;                            // previous hidden state, carried through when t >= lengths[m]
  const int32_t* lengths; int t;              // rows with t >= lengths[m] keep (h0, c0) and emit zeros
  float* seq_out; long long ld_seq_out;       // h1 (or 0 when inactive) written at seq_out[m*ld_seq_out + j]
};
struct GemmParams {
  GemmSeg seg[3];
  int nseg;
  int M, N;
  int splitk;                             // 1, 2, 4 or 8: K split over the CTAs of a cluster, reduced through DSMEM
  float* out; int ldo;                    // plain epilogue: out = act(acc + bias0 + bias1)
  const float* bias0; const float* bias1; // [N] or NULL
  const float* oscale;                    // [N] or NULL: out = act(acc + biases) * oscale[n]
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

// ---------------------------------------------------------------- pointwise.cu
// logit[b,a] = all_u_t[b,a,:] . g[b,:] + sum_d b_a[d] w_out[d] tp[b,d] + b_out   (EltwiseProdScoring rewritten)
struct ScoringParams {
  const float* all_u_t;                   // [B,A,E]
  const float* g;                         // [B,E]
  const float* tp;                        // [B,D]  linear_in_h(h_tilde) (.) w_out
  const float* b_a; const float* b_out;
  float* logit;                           // [B,A]
  int B, A, E, D;
  unsigned long long* trace;
};
int32_t launch_action_scoring(const ScoringParams& p, cudaStream_t stream);

struct TailParams {
  float* logit; const float* is_valid; const int32_t* target; int feedback; const float* sample_u;
  const float* all_u_t; int32_t* a_t; float* u_next; float* action_score; float* ce;
  int B, A, E;
  unsigned long long* trace;
};
int32_t launch_follower_tail(const TailParams& p, cudaStream_t stream);

int device_num_sms();
unsigned long long* next_trace_slot();   // NULL unless sfb_set_option("trace", 1)

}  // namespace sfb

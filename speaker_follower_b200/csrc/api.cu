// api.cu — the extern "C" entry points declared in include/sf_b200.h: argument validation, workspace
// carving and the launch sequence of one decode step.  No allocation, no synchronisation.
#include <mutex>

#include "kernels.h"

namespace sfb {

int g_disable_tc = 0;
int g_disable_fused = 0;
int g_disable_merged = 0;
int g_disable_persist = 0;
int g_disable_hoist = 0;
int g_merged_prefetch = 1;
static int g_fused_pre_weight = 4, g_fused_nopre = 0, g_fused_dbg = 0;
int g_disable_pdl = 0;
static thread_local std::string g_err;
static thread_local int g_launches = 0;

void set_error(const std::string& msg) { g_err = msg; }
void count_launch(int n) { g_launches += n; }
void reset_launch_count() { g_launches = 0; }

static int g_trace_on = 0, g_trace_n = 0;
static unsigned long long* g_trace_buf = nullptr;   // [64][16] device
unsigned long long* next_trace_slot() {
  if (!g_trace_on || !g_trace_buf || g_trace_n >= 64) return nullptr;
  return g_trace_buf + 16 * (g_trace_n++);
}

static unsigned long long* g_cta_buf = nullptr;   // [4096][8]
static int g_cta_on = 0;
int g_attn_force_cl = 0, g_attn_force_stages = 0, g_attn_no_hint = 0, g_attn_ring_kb = 0, g_attn_rb = 0;
static int g_q_with_hh = 0;
unsigned long long* cta_trace_buffer() { return g_cta_on ? g_cta_buf : nullptr; }

int device_num_sms() {   // SM count of the CURRENT device (cached per device: a process may drive several GPUs)
  static int sms[SFB_MAX_DEVICES] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  const bool tracked = dev >= 0 && dev < SFB_MAX_DEVICES;
  if (tracked && sms[dev] > 0) return sms[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  if (tracked) sms[dev] = n;
  return n;
}

// ---- workspace carving (all regions 256-byte aligned) ----
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  float* take(size_t nfloats) {
    float* r = reinterpret_cast<float*>(base + off);
    off += (nfloats * sizeof(float) + 255) & ~size_t(255);
    return r;
  }
};

struct FollowerWs {
  void* tc; size_t tc_bytes;              // tensor-core GEMM: semaphores (zeroed by the caller once) + partial tiles
  void* av; size_t av_bytes;              // visual attention: tickets + partial records
  void* at; size_t at_bytes;              // text attention
  float *tv, *q, *feat, *gates_act, *h1d, *t, *wc, *htilde, *tp, *g;
  // packed-weight path (gemm_pk.cu): semaphores + partial tiles, packed gate-GEMM activations, [t | W_out_h h] buffer
  void* pk; size_t pk_bytes;
  unsigned char* bpk; size_t bpk_bytes;
  float* th;
  void* fz; size_t fz_bytes;              // fused gather + LSTM kernel: barrier / counter words + partial tiles
  unsigned char *hdpk, *htpk;             // fused text-side kernel: packed h1 (after dropout) and packed h~ operands
  void* tsync;                            // fused text-side kernel: hand-off counters (256 B, zeroed once)
  int ldg;
  int splitk;
  size_t bytes;
};

static int kblocks(int k) { return (k + 63) / 64; }

static int gates_splitk(const sfb_dims& d, int B) {
  return gemm_pick_splitk(B, 4 * d.H, d.E + d.F + d.H, device_num_sms());
}

static FollowerWs carve_follower(const sfb_dims& d, int B, int L, int A, void* ws) {
  (void)A;
  FollowerWs w;
  Carver c(ws);
  w.splitk = gates_splitk(d, B);
  w.tc_bytes = (d.H % 32) == 0 ? gemm_tc_plan(B, d.H, d.E + d.F + d.H, 3, device_num_sms()).bytes : 256;
  w.tc = c.take(w.tc_bytes / sizeof(float));
  w.av_bytes = attention_plan(B, d.V, d.F, device_num_sms()).bytes;
  w.av = c.take(w.av_bytes / sizeof(float));
  w.at_bytes = attention_plan(B, L > 0 ? L : 1, d.H, device_num_sms()).bytes;
  w.at = c.take(w.at_bytes / sizeof(float));
  const int kmax = d.F > d.E ? d.F : d.E;
  w.tv = c.take((size_t)B * d.D);
  w.q = c.take((size_t)B * kmax);
  w.feat = c.take((size_t)B * d.F);
  w.gates_act = c.take((size_t)B * 4 * d.H);
  w.h1d = c.take((size_t)B * d.H);
  w.t = c.take((size_t)B * d.H);
  w.wc = c.take((size_t)B * d.H);
  w.htilde = c.take((size_t)B * d.H);
  w.tp = c.take((size_t)B * d.D);
  w.ldg = kmax + 4;
  w.g = c.take((size_t)B * w.ldg);
  {
    const int sms = device_num_sms();
    const int nkb_h = kblocks(d.H), nkb_g = kblocks(d.E) + kblocks(d.F) + kblocks(d.H);
    size_t mx = gemm_pk_plan(B, 4 * d.H, nkb_g, true, sms).bytes;
    const int rows[7] = {d.F, 2 * d.H, d.H, d.E + 1, 2 * d.H + d.F, ((d.E + 1 + 127) / 128) * 128 + d.F, d.H + d.F};
    for (int i = 0; i < 7; ++i) {
      const size_t b = gemm_pk_plan(B, rows[i], nkb_h, false, sms).bytes;
      if (b > mx) mx = b;
    }
    w.pk_bytes = mx;
    w.pk = c.take(mx / sizeof(float));
    w.bpk_bytes = pk_act_bytes(B, nkb_g);
    w.bpk = reinterpret_cast<unsigned char*>(c.take(w.bpk_bytes / sizeof(float)));
    w.th = c.take((size_t)B * 2 * d.H);
    const FusedPlan fp = vis_lstm_fused_plan(B, d.H, nkb_g, d.V, d.F, d.F, 0, sms);
    w.fz_bytes = fp.ok ? fp.bytes : 256;
    w.fz = c.take(w.fz_bytes / sizeof(float));
    const size_t hb = pk_act_bytes(B, nkb_h);
    w.hdpk = reinterpret_cast<unsigned char*>(c.take(hb / sizeof(float)));
    w.htpk = reinterpret_cast<unsigned char*>(c.take(hb / sizeof(float)));
    w.tsync = c.take(64);
  }
  w.bytes = c.off;
  return w;
}

// ---- state carried from one decode step to the next (opaque to the caller): the visual query of the coming step and
// the gate GEMM's packed activation operand, whose u_prev / h_0 blocks the previous step has already filled in
struct CarryLayout { size_t q, bpk, bytes; };
static CarryLayout layout_carry(const sfb_dims& d, int B) {
  CarryLayout L{};
  L.q = 0;
  L.bpk = (((size_t)B * d.F * sizeof(float)) + 255) & ~size_t(255);
  L.bytes = L.bpk + ((pk_act_bytes(B, kblocks(d.E) + kblocks(d.F) + kblocks(d.H)) + 255) & ~size_t(255));
  return L;
}

// ---- packed follower-decoder weights: offsets into the caller-owned blob
struct FollowerPk {
  size_t a_q, b_q, a_gates, a_th, a_wc, a_g, a_gq, b_g, a_kin, mq, mg, bytes;
  int nkb_h, nkb_gates;
};
static FollowerPk layout_follower_pk(const sfb_dims& d) {
  FollowerPk L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t r = off; off += (bytes + 255) & ~size_t(255); return r; };
  L.nkb_h = kblocks(d.H);
  L.nkb_gates = kblocks(d.E) + kblocks(d.F) + kblocks(d.H);
  L.a_th = take(pk_weight_bytes(2 * d.H, L.nkb_h));   // [W_in ; W_out_h] immediately followed by M_q:
  L.a_q = take(pk_weight_bytes(d.F, L.nkb_h));        // one contiguous 2H+F row operand for the fused projection
  L.b_q = take((size_t)d.F * 4);
  L.a_gates = take(pk_weight_bytes(4 * d.H, L.nkb_gates));
  L.a_kin = take(pk_weight_bytes(d.H, L.nkb_h));      // W_in^T immediately followed by W_out_c: one operand for the
  L.a_wc = take(pk_weight_bytes(d.H, L.nkb_h));       // per-episode projection [ctx W_in | ctx W_out_c^T]
  L.a_g = take(pk_weight_bytes(d.E + 1, L.nkb_h));    // M_g (+ constant row) immediately followed by a second copy
  L.a_gq = take(pk_weight_bytes(d.F, L.nkb_h));       // of M_q: one operand for the fused [g | next query] projection
  L.b_g = take((size_t)(d.E + 4) * 4);
  L.mq = take((size_t)d.F * d.H * 4);
  L.mg = take((size_t)(d.E + 1) * d.H * 4);
  L.bytes = off;
  return L;
}
static int32_t check_packable(const sfb_dims& d) {
  SFB_CHECK_ARG((d.H % 128) == 0 && (d.E % 8) == 0 && (d.F % 8) == 0 && (d.D % 4) == 0,
                "packed path needs H % 128 == 0 and E, F % 8 == 0");
  return 0;
}

struct SpkDecWs {
  void* tc; size_t tc_bytes;
  void* at; size_t at_bytes;
  void* pk; size_t pk_bytes;              // packed path: barrier words + partial tiles
  float* th;                              // packed path: [t | W_out_h h]
  float *gates_act, *h1d, *t, *wc, *htilde;
  int splitk;
  size_t bytes;
};

static SpkDecWs carve_spkdec(int H, int Ew, int B, int T, void* ws) {
  SpkDecWs w;
  Carver c(ws);
  w.splitk = gemm_pick_splitk(B, 4 * H, Ew + H, device_num_sms());
  w.tc_bytes = (H % 32) == 0 ? gemm_tc_plan(B, H, Ew + H, 2, device_num_sms()).bytes : 256;
  w.tc = c.take(w.tc_bytes / sizeof(float));
  w.at_bytes = attention_plan(B, T > 0 ? T : 1, H, device_num_sms()).bytes;
  w.at = c.take(w.at_bytes / sizeof(float));
  w.gates_act = c.take((size_t)B * 4 * H);
  w.h1d = c.take((size_t)B * H);
  w.t = c.take((size_t)B * H);
  w.wc = c.take((size_t)B * H);
  w.htilde = c.take((size_t)B * H);
  {
    const int sms = device_num_sms(), nkb_h = kblocks(H), nkb_g = kblocks(Ew) + kblocks(H);
    size_t mx = gemm_pk_plan(B, 4 * H, nkb_g, false, sms).bytes;
    const int rows[3] = {2 * H, H, 4096};   // [t|hh], h~, vocabulary projection (sized for vocab <= 4096 rows; checked at call)
    for (int i = 0; i < 3; ++i) {
      const size_t b = gemm_pk_plan(B, rows[i], nkb_h, false, sms).bytes;
      if (b > mx) mx = b;
    }
    w.pk_bytes = mx;
    w.pk = c.take(mx / sizeof(float));
    w.th = c.take((size_t)B * 2 * H);
  }
  w.bytes = c.off;
  return w;
}

struct EncWs {
  void* tc; size_t tc_bytes;
  void* pk; size_t pk_bytes;            // gemm_pk barrier words + partial tiles
  unsigned char* whh_pk[2]; size_t whh_bytes;   // W_hh packed per call (4 MB: negligible next to maxlen recurrent steps)
  unsigned char* wih_pk[2]; size_t wih_bytes;   // W_ih packed per call for the hoisted input projection
  float *xproj, *h[2], *c[2];
  unsigned char* ep_hpk; unsigned long long* ep_bar;   // persistent recurrent kernel: packed h between steps, step counters
  int splitk;
  size_t bytes;
};

static EncWs carve_encoder(int ndir, int Hd, int Ew, int B, int maxlen, void* ws) {
  EncWs w;
  Carver c(ws);
  w.splitk = gemm_pick_splitk(B, 4 * Hd, Hd, device_num_sms());
  w.tc_bytes = (Hd % 32) == 0 ? gemm_tc_plan(B, Hd, Hd, 1, device_num_sms()).bytes : 256;
  w.tc = c.take(w.tc_bytes / sizeof(float));
  w.pk_bytes = 256;
  if ((Hd % 32) == 0) {   // recurrent step (M = B) and hoisted input projection (M = B*maxlen) share the region
    const size_t a = gemm_pk_plan(B, 4 * Hd, kblocks(Hd), false, device_num_sms()).bytes;
    const size_t b = gemm_pk_plan(B * maxlen, 4 * Hd, kblocks(Ew), false, device_num_sms()).bytes;
    w.pk_bytes = a > b ? a : b;
  }
  w.wih_bytes = (Hd % 32) == 0 ? pk_weight_bytes(4 * Hd, kblocks(Ew)) : 256;
  for (int i = 0; i < 2; ++i) w.wih_pk[i] = reinterpret_cast<unsigned char*>(c.take(w.wih_bytes / sizeof(float)));
  w.pk = c.take(w.pk_bytes / sizeof(float));
  w.whh_bytes = (Hd % 32) == 0 ? pk_weight_bytes(4 * Hd, kblocks(Hd)) : 256;
  for (int i = 0; i < 2; ++i) w.whh_pk[i] = reinterpret_cast<unsigned char*>(c.take(w.whh_bytes / sizeof(float)));
  w.xproj = c.take((size_t)ndir * B * maxlen * 4 * Hd);
  for (int i = 0; i < 2; ++i) w.h[i] = c.take((size_t)ndir * B * Hd);
  for (int i = 0; i < 2; ++i) w.c[i] = c.take((size_t)ndir * B * Hd);
  const EncPersistPlan ep = encoder_persist_plan(ndir, Hd, B, device_num_sms());
  w.ep_hpk = reinterpret_cast<unsigned char*>(c.take((ep.ok ? ep.hpk_bytes : 256) / sizeof(float)));
  w.ep_bar = reinterpret_cast<unsigned long long*>(c.take((ep.ok ? ep.bar_bytes : 256) / sizeof(float)));
  w.bytes = c.off;
  return w;
}

static int32_t check_dims(const sfb_dims* d) {
  SFB_CHECK_ARG(d != nullptr, "dims is NULL");
  SFB_CHECK_ARG(d->E > 0 && d->F > 0 && d->H > 0 && d->D > 0 && d->V > 0, "dims must be positive");
  SFB_CHECK_ARG((d->E % 4) == 0 && (d->F % 4) == 0 && (d->H % 4) == 0 && (d->D % 4) == 0,
                "E, F, H, D must be multiples of 4");
  return 0;
}

static int32_t check_ws(void* ws, size_t have, size_t need) {
  if (ws == nullptr || (reinterpret_cast<uintptr_t>(ws) & 255u) != 0 || have < need) {
    set_error("workspace missing, not 256-byte aligned or too small (need " + std::to_string(need) + " bytes, have " +
              std::to_string(have) + ")");
    return SFB_ERR_WORKSPACE;
  }
  return 0;
}

// q = W_v^T (W_h h + b_h)  [B,F]; the b_v . t term is constant over the views and cancels in the softmax.
static int32_t visual_query(const sfb_dims& d, const sfb_vis_lstm_weights& w, int B, const float* h, float* tv,
                            float* q, cudaStream_t st) {
  GemmParams g{};
  g.nseg = 1;
  g.seg[0] = GemmSeg{h, d.H, nullptr, nullptr, 0, w.va_w_h, d.H, d.H, 0};
  g.M = B; g.N = d.D; g.splitk = gemm_pick_splitk(B, d.D, d.H, device_num_sms()); g.out = tv; g.ldo = d.D; g.bias0 = w.va_b_h;
  SFB_PROPAGATE(launch_gemm(g, st));
  GemmParams g2{};
  g2.nseg = 1;
  g2.seg[0] = GemmSeg{tv, d.D, nullptr, nullptr, 0, w.va_w_v, d.F, d.D, 1};
  g2.M = B; g2.N = d.F; g2.splitk = gemm_pick_splitk(B, d.F, d.D, device_num_sms()); g2.out = q; g2.ldo = d.F;
  return launch_gemm(g2, st);
}

static int32_t visual_attend(const sfb_dims& d, int B, const float* q, const sfb_visual_source& v, float* feature,
                             float* alpha_v, void* aws, size_t aws_bytes, cudaStream_t st,
                             const AttnParams* pk = nullptr) {
  AttnParams a{};
  if (pk) {
    a.pk_out = pk->pk_out; a.pk_kb0 = pk->pk_kb0; a.pk_nkb = pk->pk_nkb; a.pk_NB = pk->pk_NB;
    a.pk_rows_per_z = pk->pk_rows_per_z; a.pk_scale = pk->pk_scale; a.pk_ldscale = pk->pk_ldscale;
    a.has_side = pk->has_side; a.side = pk->side;
  }
  a.q = q; a.ldq = d.F;
  a.R = d.V; a.D = d.F;
  if (v.visual) {
    a.segA = v.visual; a.strideA_b = (long long)d.V * d.F; a.strideA_r = d.F; a.lenA = d.F;
    a.segB = nullptr; a.lenB = 0;
  } else {
    SFB_CHECK_ARG(v.feat_table && v.loc_table && v.vp_idx && v.view_idx, "gather visual source needs tables + indices");
    SFB_CHECK_ARG(v.img_dim > 0 && v.img_dim < d.F && (v.img_dim % 4) == 0, "bad img_dim");
    const int loc = d.F - v.img_dim;
    a.segA = v.feat_table; a.strideA_b = (long long)d.V * v.img_dim; a.strideA_r = v.img_dim; a.lenA = v.img_dim;
    a.idxA = v.vp_idx;
    a.idx_dependent = v.idx_dependent;
    a.segB = v.loc_table; a.strideB_b = (long long)d.V * loc; a.strideB_r = loc; a.lenB = loc;
    a.idxB = v.view_idx;
  }
  a.mask = nullptr;
  a.out = feature; a.ldo = d.F;
  a.alpha = alpha_v; a.ldalpha = d.V;
  return launch_soft_dot_attention(a, B, aws, aws_bytes, st);
}

// LSTMCell([xa | feat] .* drop_x, (h0, c0)) as ONE GEMM with the cell update fused into its epilogue.
static int32_t lstm_cell(int Ea, int F, int H, const float* w_ih, const float* w_hh, const float* b_ih,
                         const float* b_hh, int B, const float* xa, const int32_t* xa_rows, const float* feat,
                         const float* h0, const float* c0, const float* drop_x, const float* drop_h,
                         float* gates_act, float* h1, float* c1, float* h1d, void* tc, size_t tc_bytes,
                         cudaStream_t st) {
  GemmParams g{};
  const int ldw = Ea + F;
  int s = 0;
  g.seg[s++] = GemmSeg{xa, Ea, xa_rows, drop_x, drop_x ? ldw : 0, w_ih, ldw, Ea, 0};
  if (F > 0) g.seg[s++] = GemmSeg{feat, F, nullptr, drop_x ? drop_x + Ea : nullptr, drop_x ? ldw : 0, w_ih + Ea, ldw, F, 0};
  g.seg[s++] = GemmSeg{h0, H, nullptr, nullptr, 0, w_hh, H, H, 0};
  g.nseg = s;
  g.M = B; g.N = 4 * H;
  g.splitk = gemm_pick_splitk(B, 4 * H, Ea + F + H, device_num_sms());
  g.lstm.H = H; g.lstm.b_ih = b_ih; g.lstm.b_hh = b_hh; g.lstm.c0 = c0; g.lstm.drop_h = drop_h;
  g.lstm.h1 = h1; g.lstm.c1 = c1; g.lstm.h1_drop = h1d; g.lstm.gates_act = gates_act;
  if (!g_disable_tc && gemm_tc_supported(g)) return launch_gemm_tc(g, st, tc, tc_bytes);   // tcgen05 (bf16x3)
  return launch_gemm(g, st);                                                 // exact-fp32 FFMA path (any shape)
}

// SoftDotAttention on an already (optionally dropped) h: t = W_in h; attention over ctx; h~ = tanh(W_out [wc;h])
static int32_t soft_dot(int H, const sfb_softdot_weights& w, int B, int L, const float* h, const float* ctx,
                        const uint8_t* mask, float* t, float* wc, float* h_tilde, float* alpha, void* aws,
                        size_t aws_bytes, cudaStream_t st) {
  GemmParams g{};
  g.nseg = 1;
  g.seg[0] = GemmSeg{h, H, nullptr, nullptr, 0, w.w_in, H, H, 0};
  g.M = B; g.N = H; g.splitk = gemm_pick_splitk(B, H, H, device_num_sms()); g.out = t; g.ldo = H;
  SFB_PROPAGATE(launch_gemm(g, st));
  AttnParams a{};
  a.q = t; a.ldq = H; a.R = L; a.D = H;
  a.segA = ctx; a.strideA_b = (long long)L * H; a.strideA_r = H; a.lenA = H; a.lenB = 0;
  a.mask = mask; a.ldmask = L;
  a.out = wc; a.ldo = H; a.alpha = alpha; a.ldalpha = L;
  SFB_PROPAGATE(launch_soft_dot_attention(a, B, aws, aws_bytes, st));
  GemmParams g2{};
  g2.nseg = 2;
  g2.seg[0] = GemmSeg{wc, H, nullptr, nullptr, 0, w.w_out, 2 * H, H, 0};
  g2.seg[1] = GemmSeg{h, H, nullptr, nullptr, 0, w.w_out + H, 2 * H, H, 0};
  g2.M = B; g2.N = H; g2.splitk = gemm_pick_splitk(B, H, 2 * H, device_num_sms()); g2.out = h_tilde; g2.ldo = H; g2.act = 1;
  return launch_gemm(g2, st);
}

// EltwiseProdScoring.forward (model.py:342-352) re-associated: tp = w_out (.) (W_h h~ + b_h), g = W_a^T tp,
// logit_a = u_a . g + b_a . tp + b_out
static int32_t eltwise_scoring(const sfb_dims& d, const sfb_scoring_weights& wsc, int B, int A, const float* h_tilde,
                               const float* all_u_t, float* logit, float* tp, float* gbuf, cudaStream_t st) {
  GemmParams g{};
  g.nseg = 1;
  g.seg[0] = GemmSeg{h_tilde, d.H, nullptr, nullptr, 0, wsc.w_h, d.H, d.H, 0};
  g.M = B; g.N = d.D; g.splitk = gemm_pick_splitk(B, d.D, d.H, device_num_sms()); g.out = tp; g.ldo = d.D; g.bias0 = wsc.b_h; g.oscale = wsc.w_out;
  SFB_PROPAGATE(launch_gemm(g, st));
  GemmParams g2{};
  g2.nseg = 1;
  g2.seg[0] = GemmSeg{tp, d.D, nullptr, nullptr, 0, wsc.w_a, d.E, d.D, 1};
  g2.M = B; g2.N = d.E; g2.splitk = gemm_pick_splitk(B, d.E, d.D, device_num_sms()); g2.out = gbuf; g2.ldo = d.E;
  SFB_PROPAGATE(launch_gemm(g2, st));
  ScoringParams sp{};
  sp.all_u_t = all_u_t; sp.g = gbuf; sp.tp = tp; sp.b_a = wsc.b_a; sp.b_out = wsc.b_out; sp.ldg = d.E;
  sp.logit = logit; sp.B = B; sp.A = A; sp.E = d.E; sp.D = d.D;
  return launch_action_scoring(sp, st);
}

}  // namespace sfb

using namespace sfb;

extern "C" {

int32_t sfb_abi_version(void) { return SFB_ABI_VERSION; }
const char* sfb_last_error(void) { return g_err.c_str(); }
int32_t sfb_last_launch_count(void) { return g_launches; }

int32_t sfb_set_option(const char* name, int32_t value) {
  const std::string n(name ? name : "");
  if (n == "disable_tc") { g_disable_tc = value; return 0; }
  if (n == "disable_pdl") { g_disable_pdl = value; return 0; }
  if (n == "disable_fused") { g_disable_fused = value; return 0; }
  if (n == "disable_merged") { g_disable_merged = value; return 0; }
  if (n == "disable_persist") { g_disable_persist = value; return 0; }
  if (n == "disable_hoist") { g_disable_hoist = value; return 0; }
  if (n == "merged_prefetch") { g_merged_prefetch = value; return 0; }
  if (n == "fused_pre_weight") { g_fused_pre_weight = value; return 0; }
  if (n == "fused_nopre") { g_fused_nopre = value; return 0; }
  if (n == "fused_dbg") { g_fused_dbg = value; return 0; }
  if (n == "trace") {   // value 1: (re)start recording kernel slots; 0: stop
    g_trace_on = value;
    g_trace_n = 0;
    if (value && !g_trace_buf) {
      if (cudaMalloc(&g_trace_buf, 64 * 16 * sizeof(unsigned long long)) != cudaSuccess) return SFB_ERR_CUDA;
    }
    if (value) cudaMemset(g_trace_buf, 0, 64 * 16 * sizeof(unsigned long long));
    return 0;
  }
  if (n == "tc_debug") { gemm_tc_set_debug(value); return 0; }
  if (n == "attn_cl") { g_attn_force_cl = value; return 0; }
  if (n == "attn_stages") { g_attn_force_stages = value; return 0; }
  if (n == "attn_ring_kb") { g_attn_ring_kb = value; return 0; }
  if (n == "attn_rb") { g_attn_rb = value; return 0; }
  if (n == "q_with_hh") { g_q_with_hh = value; return 0; }
  if (n == "attn_nohint") { g_attn_no_hint = value; return 0; }
  if (n == "cta_trace") {   // per-CTA timeline of the attention kernel (last launch wins)
    g_cta_on = value;
    if (value && !g_cta_buf && cudaMalloc(&g_cta_buf, 4096 * 8 * sizeof(unsigned long long)) != cudaSuccess) return SFB_ERR_CUDA;
    if (value) cudaMemset(g_cta_buf, 0, 4096 * 8 * sizeof(unsigned long long));
    return 0;
  }
  set_error("unknown option: " + n);
  return SFB_ERR_INVALID_ARG;
}

int32_t sfb_debug_read_trace(int64_t* out, int32_t max_slots) {
  if (!g_trace_buf) return 0;
  const int n = g_trace_n < max_slots ? g_trace_n : max_slots;
  if (cudaMemcpy(out, g_trace_buf, (size_t)n * 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return n;
}

int32_t sfb_debug_read_cta_trace(int64_t* out, int32_t max_ctas) {
  if (!g_cta_buf) return 0;
  const int n = max_ctas < 4096 ? max_ctas : 4096;
  if (cudaMemcpy(out, g_cta_buf, (size_t)n * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return n;
}

int32_t sfb_debug_max_active_clusters(int32_t cluster, int32_t smem) { return pk_max_active_clusters(cluster, smem); }

int32_t sfb_debug_read_timestamps(int64_t* out, int32_t n) {
  return gemm_tc_read_timestamps(reinterpret_cast<long long*>(out), n);
}

int32_t sfb_device_info(int32_t* sm, int32_t* num_sms, int32_t* smem_per_block) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device");
    return SFB_ERR_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    set_error("cudaGetDeviceProperties failed");
    return SFB_ERR_NO_DEVICE;
  }
  if (sm) *sm = prop.major * 10 + prop.minor;
  if (num_sms) *num_sms = prop.multiProcessorCount;
  if (smem_per_block) *smem_per_block = (int32_t)prop.sharedMemPerBlockOptin;
  if (prop.major != 10) {
    set_error("this library carries sm_100a code only; device is sm_" + std::to_string(prop.major * 10 + prop.minor));
    return SFB_ERR_NO_DEVICE;
  }
  return 0;
}

size_t sfb_follower_step_workspace_bytes(const sfb_dims* dims, int32_t B, int32_t L, int32_t A) {
  if (!dims || B < 1) return 0;
  return carve_follower(*dims, B, L, A, nullptr).bytes;
}

size_t sfb_speaker_decoder_step_workspace_bytes(int32_t H, int32_t Ew, int32_t B, int32_t T) {
  if (H < 1 || Ew < 1 || B < 1 || T < 1) return 0;
  return carve_spkdec(H, Ew, B, T, nullptr).bytes;
}

int32_t sfb_visual_attention_fwd(const sfb_dims* dims, const sfb_vis_lstm_weights* w, int32_t B, const float* h,
                                 const sfb_visual_source* vis, float* feature, float* alpha_v, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_CHECK_ARG(w && vis && h && feature, "NULL argument");
  SFB_CHECK_ARG(B >= 1, "B >= 1");
  FollowerWs ws = carve_follower(*dims, B, 1, 1, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, ws.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SFB_PROPAGATE(visual_query(*dims, *w, B, h, ws.tv, ws.q, st));
  return visual_attend(*dims, B, ws.q, *vis, feature, alpha_v, ws.av, ws.av_bytes, st);
}

int32_t sfb_visual_attention_core_fwd(const sfb_dims* dims, int32_t B, const float* q, const sfb_visual_source* vis,
                                      float* feature, float* alpha_v, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_CHECK_ARG(q && vis && feature, "NULL argument");
  SFB_CHECK_ARG(B >= 1, "B >= 1");
  FollowerWs ws = carve_follower(*dims, B, 1, 1, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, ws.bytes));
  return visual_attend(*dims, B, q, *vis, feature, alpha_v, ws.av, ws.av_bytes, static_cast<cudaStream_t>(stream));
}

int32_t sfb_soft_dot_attention_fwd(const sfb_dims* dims, const sfb_softdot_weights* w, int32_t B, int32_t L,
                                   const float* h, const float* ctx, const uint8_t* mask, float* h_tilde,
                                   float* alpha, void* workspace, size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_CHECK_ARG(w && h && ctx && h_tilde, "NULL argument");
  SFB_CHECK_ARG(B >= 1 && L >= 1, "B, L >= 1");
  FollowerWs ws = carve_follower(*dims, B, L, 1, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, ws.bytes));
  return soft_dot(dims->H, *w, B, L, h, ctx, mask, ws.t, ws.wc, h_tilde, alpha, ws.at, ws.at_bytes,
                  static_cast<cudaStream_t>(stream));
}

int32_t sfb_follower_step_fwd(const sfb_dims* dims, const sfb_vis_lstm_weights* wl, const sfb_softdot_weights* wt,
                              const sfb_scoring_weights* wsc, int32_t B, int32_t L, int32_t A, const float* u_prev,
                              const float* all_u_t, const sfb_visual_source* vis, const float* h0, const float* c0,
                              const float* ctx, const uint8_t* ctx_mask, const float* drop_x, const float* drop_h,
                              float* h1, float* c1, float* alpha, float* logit, float* alpha_v, void* workspace,
                              size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_CHECK_ARG(wl && wt && wsc && vis, "NULL weight/source struct");
  SFB_CHECK_ARG(u_prev && all_u_t && h0 && c0 && ctx && h1 && c1 && logit, "NULL tensor argument");
  SFB_CHECK_ARG(B >= 1 && L >= 1 && A >= 1, "B, L, A >= 1");
  const sfb_dims& d = *dims;
  FollowerWs ws = carve_follower(d, B, L, A, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, ws.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  // model.py:389  feature, alpha_v = visual_attention_layer(h_0, visual_context)
  SFB_PROPAGATE(visual_query(d, *wl, B, h0, ws.tv, ws.q, st));
  SFB_PROPAGATE(visual_attend(d, B, ws.q, *vis, ws.feat, alpha_v, ws.av, ws.av_bytes, st));
  // model.py:391-394  LSTMCell(drop(cat(u_t_prev, feature)), (h_0, c_0)); h_1_drop = drop(h_1)
  SFB_PROPAGATE(lstm_cell(d.E, d.F, d.H, wl->lstm_w_ih, wl->lstm_w_hh, wl->lstm_b_ih, wl->lstm_b_hh, B, u_prev, nullptr,
                          ws.feat, h0, c0, drop_x, drop_h, ws.gates_act, h1, c1, ws.h1d, ws.tc, ws.tc_bytes, st));
  // model.py:395  h_tilde, alpha = text_attention_layer(h_1_drop, ctx, ctx_mask)
  SFB_PROPAGATE(soft_dot(d.H, *wt, B, L, ws.h1d, ctx, ctx_mask, ws.t, ws.wc, ws.htilde, alpha, ws.at, ws.at_bytes, st));
  // model.py:396  logit = decoder2action(h_tilde, all_u_t)
  SFB_PROPAGATE(eltwise_scoring(d, *wsc, B, A, ws.htilde, all_u_t, logit, ws.tp, ws.g, st));
  return 0;
}

int32_t sfb_eltwise_prod_scoring_fwd(const sfb_dims* dims, const sfb_scoring_weights* wsc, int32_t B, int32_t A,
                                     const float* h_tilde, const float* all_u_t, float* logit, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_CHECK_ARG(wsc && h_tilde && all_u_t && logit, "NULL argument");
  SFB_CHECK_ARG(B >= 1 && A >= 1, "B, A >= 1");
  FollowerWs ws = carve_follower(*dims, B, 1, A, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, ws.bytes));
  return eltwise_scoring(*dims, *wsc, B, A, h_tilde, all_u_t, logit, ws.tp, ws.g, static_cast<cudaStream_t>(stream));
}

int32_t sfb_follower_step_tail(int32_t B, int32_t A, int32_t E, float* logit, const float* is_valid,
                               const int32_t* target, int32_t feedback, const float* sample_u, const float* all_u_t,
                               int32_t* a_t, float* u_next, float* action_score, float* ce, void* stream) {
  reset_launch_count();
  SFB_CHECK_ARG(B >= 1 && A >= 1 && E >= 4 && (E % 4) == 0, "bad sizes");
  SFB_CHECK_ARG(logit && is_valid && a_t, "NULL argument");
  SFB_CHECK_ARG(feedback >= 0 && feedback <= 2, "feedback must be 0 (teacher), 1 (argmax) or 2 (sample)");
  SFB_CHECK_ARG(feedback != 0 || target, "teacher feedback needs target");
  SFB_CHECK_ARG(feedback != 2 || sample_u, "sample feedback needs sample_u");
  SFB_CHECK_ARG(!u_next || all_u_t, "u_next needs all_u_t");
  SFB_CHECK_ARG(!ce || target, "ce needs target");
  TailParams p{logit, is_valid, target, feedback, sample_u, all_u_t, a_t, u_next, action_score, ce, B, A, E};
  return launch_follower_tail(p, static_cast<cudaStream_t>(stream));
}

int32_t sfb_nav_step(const sfb_nav_tables* t, int32_t B, int32_t* state, int32_t* ended, const int32_t* goal,
                     const int32_t* a_prev, int32_t* actions_log, int32_t* vp_idx, int32_t* view_idx, int32_t* cand_view,
                     float* cand_trig, float* is_valid, int32_t* target, void* stream) {
  reset_launch_count();
  SFB_CHECK_ARG(t && t->vp && t->view && t->nvalid && t->cv && t->trig && t->next && t->A >= 1 && t->S >= 1, "nav tables missing");
  SFB_CHECK_ARG(B >= 1 && state && ended && vp_idx && view_idx && cand_view && cand_trig && is_valid, "NULL argument");
  SFB_CHECK_ARG(!target || !t->teach || goal, "teacher targets need the goal of every row");
  SFB_CHECK_ARG((reinterpret_cast<uintptr_t>(t->trig) & 15u) == 0 && (reinterpret_cast<uintptr_t>(cand_trig) & 15u) == 0, "trig tables must be 16-byte aligned");
  NavStepParams p{t->vp, t->view, t->nvalid, t->cv, t->trig, t->next, t->teach, goal, B, t->A, t->G, state, ended, a_prev, actions_log,
                  vp_idx, view_idx, cand_view, cand_trig, is_valid, target};
  return launch_nav_step(p, static_cast<cudaStream_t>(stream));
}

size_t sfb_encoder_lstm_workspace_bytes(int32_t ndir, int32_t Hd, int32_t Ew, int32_t B, int32_t maxlen) {
  if (ndir < 1 || ndir > 2 || Hd < 1 || Ew < 1 || B < 1 || maxlen < 1) return 0;
  return carve_encoder(ndir, Hd, Ew, B, maxlen, nullptr).bytes;
}

// tape of a training forward (unidirectional): activated gates of every time step and the (h, c) state before / after it
struct EncTape { float* gates; float* h; float* c; size_t bytes; };
static EncTape carve_enc_tape(int Hd, int B, int maxlen, void* p) {
  EncTape t;
  Carver c(p);
  t.gates = c.take((size_t)maxlen * B * 4 * Hd);
  t.h = c.take((size_t)(maxlen + 1) * B * Hd);
  t.c = c.take((size_t)(maxlen + 1) * B * Hd);
  t.bytes = c.off;
  return t;
}

static int32_t encoder_fwd_impl(const sfb_encoder_weights* w, int32_t ndir, int32_t Hd, int32_t Ew, int32_t B,
                                int32_t maxlen, const int32_t* seq, const int32_t* lengths, const float* drop_embed,
                                float* ctx, float* decoder_init, float* c_t, void* workspace, size_t workspace_bytes,
                                void* stream, void* tape_mem, int32_t vocab);

int32_t sfb_encoder_lstm_fwd(const sfb_encoder_weights* w, int32_t ndir, int32_t Hd, int32_t Ew, int32_t B,
                             int32_t maxlen, const int32_t* seq, const int32_t* lengths, const float* drop_embed,
                             float* ctx, float* decoder_init, float* c_t, void* workspace, size_t workspace_bytes,
                             void* stream) {
  return encoder_fwd_impl(w, ndir, Hd, Ew, B, maxlen, seq, lengths, drop_embed, ctx, decoder_init, c_t, workspace, workspace_bytes,
                          stream, nullptr, 0);
}

int32_t sfb_encoder_lstm_fwd_vocab(const sfb_encoder_weights* w, int32_t vocab, int32_t ndir, int32_t Hd, int32_t Ew, int32_t B,
                                   int32_t maxlen, const int32_t* seq, const int32_t* lengths, const float* drop_embed,
                                   float* ctx, float* decoder_init, float* c_t, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  SFB_CHECK_ARG(vocab >= 1, "vocab >= 1");
  return encoder_fwd_impl(w, ndir, Hd, Ew, B, maxlen, seq, lengths, drop_embed, ctx, decoder_init, c_t, workspace, workspace_bytes,
                          stream, nullptr, vocab);
}

size_t sfb_encoder_lstm_tape_bytes(int32_t Hd, int32_t B, int32_t maxlen) {
  if (Hd < 1 || B < 1 || maxlen < 1) return 0;
  return carve_enc_tape(Hd, B, maxlen, nullptr).bytes;
}

int32_t sfb_encoder_lstm_train_fwd(const sfb_encoder_weights* w, int32_t Hd, int32_t Ew, int32_t B, int32_t maxlen,
                                   const int32_t* seq, const int32_t* lengths, const float* drop_embed, float* ctx,
                                   float* decoder_init, float* c_t, void* tape, size_t tape_bytes, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  SFB_CHECK_ARG(tape && (reinterpret_cast<uintptr_t>(tape) & 255u) == 0 && tape_bytes >= sfb_encoder_lstm_tape_bytes(Hd, B, maxlen),
                "encoder tape missing, misaligned or too small");
  return encoder_fwd_impl(w, 1, Hd, Ew, B, maxlen, seq, lengths, drop_embed, ctx, decoder_init, c_t, workspace, workspace_bytes, stream, tape, 0);
}

static int32_t encoder_fwd_impl(const sfb_encoder_weights* w, int32_t ndir, int32_t Hd, int32_t Ew, int32_t B,
                                int32_t maxlen, const int32_t* seq, const int32_t* lengths, const float* drop_embed,
                                float* ctx, float* decoder_init, float* c_t, void* workspace, size_t workspace_bytes,
                                void* stream, void* tape_mem, int32_t vocab) {
  reset_launch_count();
  SFB_CHECK_ARG(w && seq && lengths && ctx && decoder_init && c_t, "NULL argument");
  SFB_CHECK_ARG(ndir == 1 || ndir == 2, "ndir must be 1 or 2");
  SFB_CHECK_ARG(Hd >= 4 && (Hd % 4) == 0 && Ew >= 4 && (Ew % 4) == 0 && B >= 1 && maxlen >= 1, "bad sizes");
  EncWs ws = carve_encoder(ndir, Hd, Ew, B, maxlen, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, ws.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int H = ndir * Hd;
  const size_t state = (size_t)B * Hd;
  SFB_CHECK_CUDA(cudaMemsetAsync(ws.h[0], 0, ndir * state * sizeof(float), st));
  SFB_CHECK_CUDA(cudaMemsetAsync(ws.c[0], 0, ndir * state * sizeof(float), st));
  EncTape tape{};
  const bool taped = tape_mem != nullptr;
  if (taped) {
    SFB_CHECK_ARG(ndir == 1, "the training tape covers the unidirectional encoder");
    tape = carve_enc_tape(Hd, B, maxlen, tape_mem);
    SFB_CHECK_CUDA(cudaMemsetAsync(tape.h, 0, state * sizeof(float), st));
    SFB_CHECK_CUDA(cudaMemsetAsync(tape.c, 0, state * sizeof(float), st));
    SFB_CHECK_CUDA(cudaMemsetAsync(tape.gates, 0, (size_t)maxlen * B * 4 * Hd * sizeof(float), st));   // rows past their length stay zero
  }
  int cur[2] = {0, 0};
  // recurrent projection on tcgen05 from packed W_hh (gemm_pk.cu): pack once per call, then one launch per time step
  const bool use_pk = !g_disable_tc && (Hd % 32) == 0 && (Hd % 8) == 0 &&
                      (reinterpret_cast<uintptr_t>(w->w_hh[0]) & 15u) == 0 && (ndir == 1 || (reinterpret_cast<uintptr_t>(w->w_hh[1]) & 15u) == 0);
  if (use_pk) {
    for (int dir = 0; dir < ndir; ++dir) {
      PackParams p{};
      p.nseg = 1;
      p.seg[0] = PackSeg{w->w_hh[dir], Hd, Hd, nullptr, 0, nullptr};
      p.ntile = Hd / 32; p.R = 128; p.rows_per_tile = 128; p.rows_valid = 4 * Hd; p.lstm_H = Hd;
      p.out = ws.whh_pk[dir];
      SFB_PROPAGATE(launch_pack_rows(p, st));
      PackParams pi{};   // W_ih in natural row order: the projection's columns stay [i | f | g | o]
      pi.nseg = 1;
      pi.seg[0] = PackSeg{w->w_ih[dir], Ew, Ew, nullptr, 0, nullptr};
      pi.ntile = (4 * Hd + 127) / 128; pi.R = 128; pi.rows_per_tile = 128; pi.rows_valid = 4 * Hd; pi.lstm_H = 0;
      pi.out = ws.wih_pk[dir];
      SFB_PROPAGATE(launch_pack_rows(pi, st));
    }
  }
  // the recurrence as ONE launch with W_hh resident in shared memory (encoder_persist.cu) when the shape fits
  const bool persist = use_pk && !g_disable_persist && encoder_persist_plan(ndir, Hd, B, device_num_sms()).ok;
  // Without dropout on the embedding the input projection depends on the word id only: project the vocab rows of the
  // embedding table once (T = Emb W_ih^T, [vocab, 4Hd]) and let the recurrent kernel pick rows by word id, instead of
  // projecting B * maxlen gathered rows
  const bool by_token = persist && drop_embed == nullptr && vocab > 0 && vocab <= B * maxlen;
  for (int dir = 0; dir < ndir; ++dir) {
    // hoisted input projection for every time step at once: [B*maxlen, Ew] x W_ih^T  (model.py:85,90)
    float* xp = ws.xproj + (size_t)dir * B * maxlen * 4 * Hd;
    GemmParams g{};
    g.nseg = 1;
    g.seg[0] = GemmSeg{w->embedding, Ew, by_token ? nullptr : seq, drop_embed, drop_embed ? Ew : 0, w->w_ih[dir], Ew, Ew, 0};
    g.M = by_token ? vocab : B * maxlen; g.N = 4 * Hd; g.splitk = 1; g.out = xp; g.ldo = 4 * Hd;
    if (use_pk && (reinterpret_cast<uintptr_t>(w->w_ih[dir]) & 15u) == 0 && (reinterpret_cast<uintptr_t>(w->embedding) & 15u) == 0) {
      PkParams q{};   // tcgen05, weights bulk-copied, embedding rows gathered + split on the fly
      q.g = g;
      q.a_pk = ws.wih_pk[dir]; q.b_pk = nullptr; q.nkb = kblocks(Ew);
      SFB_PROPAGATE(launch_gemm_pk(q, st, ws.pk, ws.pk_bytes));
    } else {
      SFB_PROPAGATE(launch_gemm(g, st));
    }
    for (int s = 0; s < maxlen && !persist; ++s) {
      const int t = dir == 0 ? s : maxlen - 1 - s;
      float* hp = taped ? tape.h + (size_t)s * state : ws.h[cur[dir]] + dir * state;
      float* cp = taped ? tape.c + (size_t)s * state : ws.c[cur[dir]] + dir * state;
      float* hn = taped ? tape.h + (size_t)(s + 1) * state : ws.h[cur[dir] ^ 1] + dir * state;
      float* cn = taped ? tape.c + (size_t)(s + 1) * state : ws.c[cur[dir] ^ 1] + dir * state;
      GemmParams r{};
      r.nseg = 1;
      r.seg[0] = GemmSeg{hp, Hd, nullptr, nullptr, 0, w->w_hh[dir], Hd, Hd, 0};
      r.M = B; r.N = 4 * Hd; r.splitk = ws.splitk;
      LstmEpilogue& p = r.lstm;
      p.H = Hd; p.b_ih = w->b_ih[dir]; p.b_hh = w->b_hh[dir];
      p.c0 = cp; p.h0 = hp; p.h1 = hn; p.c1 = cn;
      p.addend = xp + (size_t)t * 4 * Hd; p.ld_addend = (long long)maxlen * 4 * Hd;
      p.lengths = lengths; p.t = t;
      p.seq_out = ctx + (size_t)t * H + dir * Hd; p.ld_seq_out = (long long)maxlen * H;
      if (taped) p.gates_act = tape.gates + (size_t)s * B * 4 * Hd;
      if (use_pk) {
        PkParams q{};
        q.g = r;
        q.a_pk = ws.whh_pk[dir]; q.b_pk = nullptr; q.nkb = kblocks(Hd);
        SFB_PROPAGATE(launch_gemm_pk(q, st, ws.pk, ws.pk_bytes));
      } else if (!g_disable_tc && gemm_tc_supported(r)) {
        SFB_PROPAGATE(launch_gemm_tc(r, st, ws.tc, ws.tc_bytes));
      } else {
        SFB_PROPAGATE(launch_gemm(r, st));
      }
      cur[dir] ^= 1;
    }
  }
  if (persist) {
    EncPersistParams ep{};
    for (int dir = 0; dir < ndir; ++dir) {
      ep.whh_pk[dir] = ws.whh_pk[dir]; ep.xproj[dir] = ws.xproj + (size_t)dir * B * maxlen * 4 * Hd;
      ep.b_ih[dir] = w->b_ih[dir]; ep.b_hh[dir] = w->b_hh[dir];
    }
    ep.seq = by_token ? seq : nullptr;
    ep.lengths = lengths; ep.ctx = ctx; ep.ld_ctx = (long long)maxlen * H; ep.H = H;
    ep.h_fin = ws.h[0]; ep.c_fin = ws.c[0];
    if (taped) { ep.tape_h = tape.h; ep.tape_c = tape.c; ep.tape_g = tape.gates; }
    ep.hpk = ws.ep_hpk; ep.bar = ws.ep_bar;
    ep.ndir = ndir; ep.Hd = Hd; ep.B = B; ep.maxlen = maxlen;
    SFB_PROPAGATE(launch_encoder_persist(ep, st));
    cur[0] = cur[1] = 0;
  }
  // decoder_init = tanh(encoder2decoder(h_t)), h_t = cat(reverse, forward) when bidirectional (model.py:92-99)
  GemmParams e{};
  const float* h_last = taped ? tape.h + (size_t)maxlen * state : ws.h[cur[0]];
  const float* c_last = taped ? tape.c + (size_t)maxlen * state : ws.c[cur[0]];
  if (ndir == 1) {
    e.nseg = 1;
    e.seg[0] = GemmSeg{h_last, Hd, nullptr, nullptr, 0, w->e2d_w, H, Hd, 0};
  } else {
    e.nseg = 2;
    e.seg[0] = GemmSeg{ws.h[cur[1]] + state, Hd, nullptr, nullptr, 0, w->e2d_w, H, Hd, 0};
    e.seg[1] = GemmSeg{ws.h[cur[0]], Hd, nullptr, nullptr, 0, w->e2d_w + Hd, H, Hd, 0};
  }
  e.M = B; e.N = H; e.splitk = gemm_pick_splitk(B, H, H, device_num_sms()); e.out = decoder_init; e.ldo = H; e.bias0 = w->e2d_b; e.act = 1;
  SFB_PROPAGATE(launch_gemm(e, st));
  if (ndir == 1) {
    SFB_CHECK_CUDA(cudaMemcpyAsync(c_t, c_last, state * sizeof(float), cudaMemcpyDeviceToDevice, st));
  } else {
    SFB_CHECK_CUDA(cudaMemcpy2DAsync(c_t, (size_t)H * 4, ws.c[cur[1]] + state, (size_t)Hd * 4, (size_t)Hd * 4, B,
                                     cudaMemcpyDeviceToDevice, st));
    SFB_CHECK_CUDA(cudaMemcpy2DAsync(c_t + Hd, (size_t)H * 4, ws.c[cur[0]], (size_t)Hd * 4, (size_t)Hd * 4, B,
                                     cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

int32_t sfb_speaker_encoder_step_fwd(const sfb_dims* dims, const sfb_vis_lstm_weights* w, int32_t B,
                                     const float* action_embedding, const sfb_visual_source* vis, const float* h0,
                                     const float* c0, const float* drop_x, float* h1, float* c1, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_CHECK_ARG(w && vis && action_embedding && h0 && c0 && h1 && c1, "NULL argument");
  SFB_CHECK_ARG(B >= 1, "B >= 1");
  const sfb_dims& d = *dims;
  FollowerWs ws = carve_follower(d, B, 1, 1, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, ws.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SFB_PROPAGATE(visual_query(d, *w, B, h0, ws.tv, ws.q, st));
  // the attention weights stay in the workspace (tp region, [B,V] <= [B,D]) for sfb_speaker_encoder_step_bwd
  SFB_PROPAGATE(visual_attend(d, B, ws.q, *vis, ws.feat, d.V <= d.D ? ws.tp : nullptr, ws.av, ws.av_bytes, st));
  return lstm_cell(d.E, d.F, d.H, w->lstm_w_ih, w->lstm_w_hh, w->lstm_b_ih, w->lstm_b_hh, B, action_embedding, nullptr,
                   ws.feat, h0, c0, drop_x, nullptr, ws.gates_act, h1, c1, nullptr, ws.tc, ws.tc_bytes, st);
}

int32_t sfb_speaker_decoder_step_fwd(const sfb_speaker_decoder_weights* w, int32_t H, int32_t Ew, int32_t vocab,
                                     int32_t B, int32_t T, const int32_t* prev_word, const float* h0, const float* c0,
                                     const float* ctx, const uint8_t* ctx_mask, const float* drop_e,
                                     const float* drop_h, float* h1, float* c1, float* alpha, float* logit,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_CHECK_ARG(w && prev_word && h0 && c0 && ctx && h1 && c1 && logit, "NULL argument");
  SFB_CHECK_ARG(H >= 4 && (H % 4) == 0 && Ew >= 4 && (Ew % 4) == 0 && vocab >= 1 && B >= 1 && T >= 1, "bad sizes");
  SpkDecWs ws = carve_spkdec(H, Ew, B, T, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, ws.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // model.py:497-503,515: LSTMCell(embedding(previous_word)) — the lookup is a row indirection of the A operand
  SFB_PROPAGATE(lstm_cell(Ew, 0, H, w->lstm_w_ih, w->lstm_w_hh, w->lstm_b_ih, w->lstm_b_hh, B, w->embedding, prev_word,
                          nullptr, h0, c0, drop_e, drop_h, ws.gates_act, h1, c1, ws.h1d, ws.tc, ws.tc_bytes, st));
  // model.py:516-517
  SFB_PROPAGATE(soft_dot(H, w->attn, B, T, ws.h1d, ctx, ctx_mask, ws.t, ws.wc, ws.htilde, alpha, ws.at, ws.at_bytes, st));
  // model.py:518  logit = decoder2action(h_tilde)
  GemmParams g{};
  g.nseg = 1;
  g.seg[0] = GemmSeg{ws.htilde, H, nullptr, nullptr, 0, w->w_voc, H, H, 0};
  g.M = B; g.N = vocab; g.splitk = gemm_pick_splitk(B, vocab, H, device_num_sms()); g.out = logit; g.ldo = vocab; g.bias0 = w->b_voc;
  return launch_gemm(g, st);
}

size_t sfb_follower_carry_bytes(const sfb_dims* dims, int32_t B) {
  if (!dims || check_dims(dims) != 0 || check_packable(*dims) != 0 || B < 1) return 0;
  return layout_carry(*dims, B).bytes;
}

size_t sfb_follower_packed_bytes(const sfb_dims* dims) {
  if (!dims || check_dims(dims) != 0 || check_packable(*dims) != 0) return 0;
  return layout_follower_pk(*dims).bytes;
}

int32_t sfb_follower_pack_weights(const sfb_dims* dims, const sfb_vis_lstm_weights* wl, const sfb_softdot_weights* wt,
                                  const sfb_scoring_weights* wsc, void* packed, size_t packed_bytes, void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_PROPAGATE(check_packable(*dims));
  SFB_CHECK_ARG(wl && wt && wsc && packed, "NULL argument");
  const sfb_dims& d = *dims;
  const FollowerPk L = layout_follower_pk(d);
  SFB_CHECK_ARG(packed_bytes >= L.bytes && (reinterpret_cast<uintptr_t>(packed) & 255u) == 0, "packed buffer too small / misaligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* base = static_cast<unsigned char*>(packed);
  float* mq = reinterpret_cast<float*>(base + L.mq);
  float* mg = reinterpret_cast<float*>(base + L.mg);
  float* bq = reinterpret_cast<float*>(base + L.b_q);
  float* bg = reinterpret_cast<float*>(base + L.b_g);
  // M_q = W_v^T W_h, b_q = W_v^T b_h           (model.py:316-320; b_v . t cancels in the softmax)
  SFB_PROPAGATE(launch_fold(wl->va_w_v, d.F, nullptr, wl->va_w_h, d.H, wl->va_b_h, d.D, d.F, d.H, mq, d.H, bq, nullptr, st));
  // M_g = W_a^T diag(w_o) W_h', b_g = W_a^T (w_o . b_h'); extra row E carries the per-row constant
  // (b_a . w_o)^T (W_h' h~ + b_h') + b_o       (model.py:348-351)
  SFB_PROPAGATE(launch_fold(wsc->w_a, d.E, wsc->w_out, wsc->w_h, d.H, wsc->b_h, d.D, d.E, d.H, mg, d.H, bg, nullptr, st));
  SFB_PROPAGATE(launch_fold(wsc->b_a, 1, wsc->w_out, wsc->w_h, d.H, wsc->b_h, d.D, 1, d.H, mg + (size_t)d.E * d.H, d.H,
                            bg + d.E, wsc->b_out, st));
  auto plain = [&](const float* w, int ldw, int rows, int k, unsigned char* out) {
    PackParams p{};
    p.nseg = 1;
    p.seg[0] = PackSeg{w, ldw, k, nullptr, 0, nullptr};
    p.ntile = (rows + 127) / 128; p.R = 128; p.rows_per_tile = 128; p.rows_valid = rows; p.lstm_H = 0;
    p.out = out;
    return launch_pack_rows(p, st);
  };
  SFB_PROPAGATE(plain(mq, d.H, d.F, d.H, base + L.a_q));
  SFB_PROPAGATE(plain(mq, d.H, d.F, d.H, base + L.a_gq));
  {
    PackParams p{};
    p.nseg = 3;
    p.seg[0] = PackSeg{wl->lstm_w_ih, d.E + d.F, d.E, nullptr, 0, nullptr};
    p.seg[1] = PackSeg{wl->lstm_w_ih + d.E, d.E + d.F, d.F, nullptr, 0, nullptr};
    p.seg[2] = PackSeg{wl->lstm_w_hh, d.H, d.H, nullptr, 0, nullptr};
    p.ntile = d.H / 32; p.R = 128; p.rows_per_tile = 128; p.rows_valid = 4 * d.H; p.lstm_H = d.H;
    p.out = base + L.a_gates;
    SFB_PROPAGATE(launch_pack_rows(p, st));
  }
  // rows 0..H-1: W_in (t = W_in h); rows H..2H-1: W_out[:, H:2H] (the h part of linear_out)   (model.py:129,140-142)
  SFB_PROPAGATE(plain(wt->w_in, d.H, d.H, d.H, base + L.a_th));
  SFB_PROPAGATE(plain(wt->w_out + d.H, 2 * d.H, d.H, d.H, base + L.a_th + pk_weight_bytes(d.H, L.nkb_h)));
  SFB_PROPAGATE(plain(wt->w_out, 2 * d.H, d.H, d.H, base + L.a_wc));
  SFB_PROPAGATE(plain(mg, d.H, d.E + 1, d.H, base + L.a_g));
  // W_in^T (scratch: the M_q area, already consumed above) for sfb_follower_project_ctx
  SFB_PROPAGATE(launch_transpose(wt->w_in, d.H, d.H, d.H, mq, d.H, st));
  SFB_PROPAGATE(plain(mq, d.H, d.H, d.H, base + L.a_kin));
  return 0;
}

// workspace of the per-episode projection: [barrier words + partial tiles | packed ctx rows].  Bounds that hold for
// every row count 1..B*L (the caller may pass a compacted subset): split-K only happens while all CTAs are
// co-resident (<= #SMs tiles of <= 128 x 256 floats); the packed rows take <= 64 KB per K block per 128 rows (+ rounding).
static size_t project_ctx_region_bytes() { return 4096 + (size_t)device_num_sms() * 128 * 256 * sizeof(float); }

size_t sfb_follower_project_ctx_workspace_bytes(const sfb_dims* dims, int32_t B, int32_t L) {
  if (!dims || B < 1 || L < 1) return 0;
  const size_t act = ((size_t)B * L / 128 + 2) * (size_t)kblocks(dims->H) * 65536;
  return project_ctx_region_bytes() + act;
}

int32_t sfb_follower_project_ctx(const sfb_dims* dims, const void* packed, size_t packed_bytes, int32_t B, int32_t L,
                                 const float* ctx, const int32_t* rows, int32_t n_rows, float* ctx_k, float* ctx_o,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_PROPAGATE(check_packable(*dims));
  SFB_CHECK_ARG(packed && ctx && ctx_k && ctx_o && B >= 1 && L >= 1, "NULL / bad argument");
  const sfb_dims& d = *dims;
  const FollowerPk P = layout_follower_pk(d);
  SFB_CHECK_ARG(P.a_wc == P.a_kin + pk_weight_bytes(d.H, P.nkb_h), "packed layout: ctx projection operand not contiguous");
  SFB_CHECK_ARG(packed_bytes >= P.bytes && (reinterpret_cast<uintptr_t>(packed) & 255u) == 0, "packed buffer too small / misaligned");
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, sfb_follower_project_ctx_workspace_bytes(dims, B, L)));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned char* base = static_cast<const unsigned char*>(packed);
  SFB_CHECK_ARG(rows == nullptr || (n_rows >= 1 && n_rows <= B * L), "rows: 1 <= n_rows <= B*L");
  const int M = rows ? n_rows : B * L, nkb = kblocks(d.H);
  const PkPlan pl = gemm_pk_plan(M, 2 * d.H, nkb, true, device_num_sms(), true);
  const size_t region = project_ctx_region_bytes();
  SFB_CHECK_ARG((pl.S == 1 ? pl.sem_bytes : pl.bytes) <= region && region + pk_act_bytes(M, nkb, true) <= workspace_bytes,
                "project_ctx: workspace too small for this row count");
  unsigned char* xpk = static_cast<unsigned char*>(workspace) + region;
  // 1. ctx rows -> bf16 hi/lo operand tiles, once (both projections and all four weight tiles of each share them)
  PackParams pp{};
  pp.nseg = 1;
  pp.seg[0] = PackSeg{ctx, d.H, d.H, nullptr, 0, rows};   // rows: only the un-padded (b, l) positions are projected
  pp.ntile = pl.nz; pp.R = pl.NB; pp.rows_per_tile = pl.rows_per_z; pp.rows_valid = M; pp.lstm_H = 0;
  pp.out = xpk;
  SFB_PROPAGATE(launch_pack_rows(pp, st));
  // 2. [ctx_k | ctx_o] = ctx [W_in | W_out_c^T]: one tcgen05 projection, bulk-copy fed on both operands
  PkParams q{};
  q.a_pk = base + P.a_kin; q.b_pk = xpk; q.nkb = nkb;
  q.g.M = M; q.g.N = 2 * d.H; q.g.out = ctx_k; q.g.ldo = d.H;
  q.g.n_split = d.H; q.g.out2 = ctx_o; q.g.ldo2 = d.H;
  q.g.out_rows = rows;
  q.wide = 1;
  return launch_gemm_pk(q, st, workspace, region);
}

int32_t sfb_follower_step_packed_fwd(const sfb_dims* dims, const sfb_vis_lstm_weights* wl, const void* packed,
                                     size_t packed_bytes, int32_t B, int32_t L, int32_t A, const float* u_prev,
                                     const float* all_u_t, const sfb_visual_source* vis, const float* h0,
                                     const float* c0, const float* ctx, const uint8_t* ctx_mask, const float* drop_x,
                                     const float* drop_h, float* h1, float* c1, float* alpha, float* logit,
                                     float* alpha_v, void* carry_in, void* carry_out, const sfb_step_tail* tail,
                                     const sfb_action_source* act, const float* ctx_k, const float* ctx_o,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_PROPAGATE(check_packable(*dims));
  SFB_CHECK_ARG(wl && vis && packed, "NULL weight/source struct");
  SFB_CHECK_ARG(u_prev && h0 && c0 && ctx && h1 && c1 && logit, "NULL tensor argument");
  SFB_CHECK_ARG(B >= 1 && L >= 1 && A >= 1, "B, L, A >= 1");
  SFB_CHECK_ARG((ctx_k == nullptr) == (ctx_o == nullptr), "ctx_k and ctx_o must be given together");
  SFB_CHECK_ARG((!carry_in || (reinterpret_cast<uintptr_t>(carry_in) & 255u) == 0) && (!carry_out || (reinterpret_cast<uintptr_t>(carry_out) & 255u) == 0) &&
                    (!carry_in || carry_in != carry_out), "carry buffers must be 256-byte aligned and distinct");
  const bool act_gather = act && act->all_u_t == nullptr;
  if (act_gather) {
    SFB_CHECK_ARG(act->feat_table && act->vp_idx && act->cand_view && act->cand_trig, "gather action source needs table, indices and trig values");
    SFB_CHECK_ARG(act->img_dim > 0 && act->img_dim < dims->E && (act->img_dim % 4) == 0, "bad action img_dim");
  } else {
    if (act) all_u_t = act->all_u_t;
    SFB_CHECK_ARG(all_u_t, "all_u_t is NULL and no gather action source given");
  }
  if (tail) {
    SFB_CHECK_ARG(tail->is_valid && tail->a_t, "tail: is_valid and a_t are required");
    SFB_CHECK_ARG(tail->feedback >= 0 && tail->feedback <= 2, "tail: feedback must be 0 (teacher), 1 (argmax) or 2 (sample)");
    SFB_CHECK_ARG(tail->feedback != 0 || tail->target, "tail: teacher feedback needs target");
    SFB_CHECK_ARG(tail->feedback != 2 || tail->sample_u, "tail: sample feedback needs sample_u");
    SFB_CHECK_ARG(!tail->ce || tail->target, "tail: ce needs target");
  }
  const sfb_dims& d = *dims;
  const FollowerPk P = layout_follower_pk(d);
  SFB_CHECK_ARG(P.a_q == P.a_th + pk_weight_bytes(2 * d.H, P.nkb_h), "packed layout: fused projection operand not contiguous");
  SFB_CHECK_ARG(packed_bytes >= P.bytes && (reinterpret_cast<uintptr_t>(packed) & 255u) == 0, "packed buffer too small / misaligned");
  FollowerWs ws = carve_follower(d, B, L, A, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, ws.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned char* base = static_cast<const unsigned char*>(packed);
  const float* b_q = reinterpret_cast<const float*>(base + P.b_q);
  auto proj = [&](const unsigned char* a_pk, const float* x, int ldx, int k, int n_out, float* out, int ldo,
                  const float* bias, const float* padd, int ld_padd, int act, int n_split, float* out2, int ldo2,
                  const float* bias2) {
    PkParams q{};
    q.a_pk = a_pk; q.b_pk = nullptr; q.nkb = kblocks(k);
    q.g.nseg = 1;
    q.g.seg[0] = GemmSeg{x, ldx, nullptr, nullptr, 0, nullptr, 0, k, 0};
    q.g.M = B; q.g.N = n_out; q.g.out = out; q.g.ldo = ldo; q.g.bias0 = bias; q.g.padd = padd; q.g.ld_padd = ld_padd; q.g.act = act;
    q.g.n_split = n_split; q.g.out2 = out2; q.g.ldo2 = ldo2; q.g.bias2 = bias2;
    return launch_gemm_pk(q, st, ws.pk, ws.pk_bytes);
  };
  // ---- state carried across steps (see layout_carry): q of this step + the packed [u_prev | . | h_0] operand blocks.
  // Train mode (drop_x) re-packs u_prev under this step's mask, so only eval steps consume a carried operand.
  const CarryLayout CL = layout_carry(d, B);
  const PkPlan gpl = gemm_pk_plan(B, 4 * d.H, P.nkb_gates, true, device_num_sms());
  const int vis_lenA = vis->visual ? d.F : vis->img_dim, vis_lenB = vis->visual ? 0 : d.F - vis->img_dim;
  if (!vis->visual) {
    SFB_CHECK_ARG(vis->feat_table && vis->loc_table && vis->vp_idx && vis->view_idx, "gather visual source needs tables + indices");
    SFB_CHECK_ARG(vis->img_dim > 0 && vis->img_dim < d.F && (vis->img_dim % 4) == 0, "bad img_dim");
  }
  const FusedPlan fpl = vis_lstm_fused_plan(B, d.H, P.nkb_gates, d.V, d.F, vis_lenA, vis_lenB, device_num_sms());
  const bool fused = fpl.ok && !g_disable_fused && gpl.nz == 1 && gpl.NB == fpl.NB;
  const bool use_carry = carry_in != nullptr && drop_x == nullptr;
  float* q_next = carry_out ? reinterpret_cast<float*>(static_cast<char*>(carry_out) + CL.q) : nullptr;
  unsigned char* bpk_next = carry_out ? reinterpret_cast<unsigned char*>(carry_out) + CL.bpk : nullptr;
  const float* qv = use_carry ? reinterpret_cast<const float*>(static_cast<char*>(carry_in) + CL.q) : nullptr;
  unsigned char* bpk_cur = (use_carry && fused) ? reinterpret_cast<unsigned char*>(carry_in) + CL.bpk : ws.bpk;
  auto side_pack = [&](PackParams& side) {   // [u_prev (.) drop | (attention output: written elsewhere) | h_0] -> bpk_cur
    side.nseg = 3;
    side.seg[0] = PackSeg{u_prev, d.E, d.E, drop_x, drop_x ? d.E + d.F : 0, nullptr};
    side.seg[1] = PackSeg{nullptr, d.F, d.F, nullptr, 0, nullptr};
    side.seg[2] = PackSeg{h0, d.H, d.H, nullptr, 0, nullptr};
    side.ntile = gpl.nz; side.R = gpl.NB; side.rows_per_tile = gpl.rows_per_z; side.rows_valid = B; side.lstm_H = 0;
    side.out = bpk_cur;
  };
  // model.py:389  visual query q = W_v^T (W_h h_0 + b_h) = M_q h_0 + b_q: carried over from the previous step when the
  // caller passes its carry, computed here otherwise (the fused path packs u_prev / h_0 as a side job of this launch)
  if (!qv) {
    PkParams q{};
    q.a_pk = base + P.a_q; q.b_pk = nullptr; q.nkb = kblocks(d.H);
    q.g.nseg = 1;
    q.g.seg[0] = GemmSeg{h0, d.H, nullptr, nullptr, 0, nullptr, 0, d.H, 0};
    q.g.M = B; q.g.N = d.F; q.g.out = ws.q; q.g.ldo = d.F; q.g.bias0 = b_q;
    if (fused) {
      q.has_side = 1;
      side_pack(q.side);
    }
    SFB_PROPAGATE(launch_gemm_pk(q, st, ws.pk, ws.pk_bytes));
    qv = ws.q;
  }
  LstmEpilogue lstm_e{};
  lstm_e.H = d.H; lstm_e.b_ih = wl->lstm_b_ih; lstm_e.b_hh = wl->lstm_b_hh; lstm_e.c0 = c0; lstm_e.drop_h = drop_h;
  lstm_e.h1 = h1; lstm_e.c1 = c1; lstm_e.h1_drop = ws.h1d; lstm_e.gates_act = ws.gates_act;
  lstm_e.hpk_NB = gpl.NB; lstm_e.hpk_rows_per_z = gpl.rows_per_z;
  if (bpk_next && gpl.nz == 1) {   // h_1 in packed form: the h_0 blocks of the next step's gate GEMM
    lstm_e.hpk = bpk_next; lstm_e.hpk_kb0 = kblocks(d.E) + kblocks(d.F); lstm_e.hpk_nkb = P.nkb_gates;
  }
  // the text side as ONE launch (step_fused_b.cu) when the per-episode ctx projections are given and the shape fits
  const FusedTextPlan tpl = text_score_fused_plan(B, L, A, d.H, d.E, d.F, q_next != nullptr, device_num_sms());
  const bool fused_b = ctx_k != nullptr && tpl.ok && !g_disable_fused && gpl.nz == 1 && gpl.NB == tpl.NB &&
                       (!act_gather || (size_t)A * d.E * 4 <= 160 * 1024);
  if (fused_b) lstm_e.hdpk = ws.hdpk;
  // ... and the WHOLE step as one launch (step_fused_b.cu: step_kernel) when both halves are fused and fit one geometry
  const bool merged = fused && fused_b && !g_disable_merged &&
                      step_fused_plan(B, L, A, d.H, d.E, d.F, q_next != nullptr, P.nkb_gates, d.V, d.F, vis_lenA, vis_lenB, device_num_sms()).ok;
  FusedVisLstmParams f{};
  if (fused) {
    // model.py:389-393 as ONE launch (step_fused.cu): attention gather + gate GEMM + LSTM cell
    f.q = qv; f.ldq = d.F; f.R = d.V; f.D = d.F;
    if (vis->visual) {
      f.segA = vis->visual; f.strideA_b = (long long)d.V * d.F; f.lenA = d.F; f.lenB = 0;
    } else {
      const int loc = d.F - vis->img_dim;
      f.segA = vis->feat_table; f.strideA_b = (long long)d.V * vis->img_dim; f.lenA = vis->img_dim; f.idxA = vis->vp_idx;
      f.segB = vis->loc_table; f.strideB_b = (long long)d.V * loc; f.lenB = loc; f.idxB = vis->view_idx;
    }
    f.feat = ws.feat; f.ldfeat = d.F; f.alpha = alpha_v; f.ldalpha = d.V;
    f.pk_scale = drop_x ? drop_x + d.E : nullptr; f.pk_ldscale = d.E + d.F;
    f.a_pk = base + P.a_gates; f.b_pk = bpk_cur; f.nkb = P.nkb_gates;
    f.post_kb0 = kblocks(d.E); f.post_kb1 = kblocks(d.E) + kblocks(d.F); f.feat_kb0 = kblocks(d.E);
    f.g.lstm = lstm_e; f.g.M = B; f.g.N = 4 * d.H;
    f.B = B;
    f.idx_dependent = vis->idx_dependent;
    f.pre_weight_free = g_fused_pre_weight;
    f.dbg = g_fused_dbg;
    if (g_fused_nopre) { f.post_kb0 = 0; f.post_kb1 = P.nkb_gates; }   // bring-up: no overlap of the gate GEMM with the gather
    if (!merged) SFB_PROPAGATE(launch_vis_lstm_fused(f, st, ws.fz, ws.fz_bytes));
  } else {
    // The gate GEMM's activation operand [u_prev | feature | h0] (.) drop_x is packed by the attention kernel: u_prev and
    // h0 as a side job of all its threads, feature in its epilogue.
    {
      AttnParams pk{};
      pk.pk_out = ws.bpk; pk.pk_kb0 = kblocks(d.E); pk.pk_nkb = P.nkb_gates; pk.pk_NB = gpl.NB; pk.pk_rows_per_z = gpl.rows_per_z;
      pk.pk_scale = drop_x ? drop_x + d.E : nullptr; pk.pk_ldscale = d.E + d.F;
      pk.has_side = 1;
      side_pack(pk.side);
      SFB_PROPAGATE(visual_attend(d, B, qv, *vis, ws.feat, alpha_v, ws.av, ws.av_bytes, st, &pk));
    }
    // model.py:391-393  LSTMCell(drop(cat(u_t_prev, feature)), (h_0, c_0)) on tcgen05 from the packed operands
    {
      PkParams q{};
      q.a_pk = base + P.a_gates; q.b_pk = ws.bpk; q.nkb = P.nkb_gates;
      q.g.M = B; q.g.N = 4 * d.H;
      q.g.lstm = lstm_e;
      SFB_PROPAGATE(launch_gemm_pk(q, st, ws.pk, ws.pk_bytes));
    }
  }
  const float* b_g = reinterpret_cast<const float*>(base + P.b_g);
  if (fused_b) {
    FusedTextScoreParams t{};
    t.B = B; t.L = L; t.A = A; t.H = d.H; t.E = d.E;
    t.h1d = ws.h1d; t.ldh = d.H; t.ctx_k = ctx_k; t.ctx_o = ctx_o; t.mask = ctx_mask; t.ldmask = L;
    t.alpha = alpha; t.ldalpha = L; t.h_tilde = ws.htilde; t.htpk = ws.htpk;
    t.a_hh = base + P.a_th + pk_weight_bytes(d.H, P.nkb_h); t.hh_tiles = d.H / 128; t.hdpk = ws.hdpk; t.hh = ws.th; t.ldhh = d.H;
    if (q_next) {
      const size_t b_half = (size_t)(gpl.NB / 8) * 1024;
      t.a_q = base + P.a_q; t.q_tiles = (d.F + 127) / 128; t.hpk = bpk_next + (size_t)(kblocks(d.E) + kblocks(d.F)) * 2 * b_half;
      t.q_next = q_next; t.ldq = d.F; t.b_q = b_q; t.q_cols = d.F;
    }
    t.a_g = base + P.a_g; t.g_tiles = (d.E + 1 + 127) / 128; t.g = ws.g; t.ldg = ws.ldg; t.b_g = b_g; t.g_cols = d.E + 1;
    t.all_u_t = all_u_t;
    if (act_gather) {
      t.all_u_t = nullptr; t.cand_table = act->feat_table; t.vp_idx = act->vp_idx; t.cand_view = act->cand_view;
      t.cand_trig = act->cand_trig; t.img_dim = act->img_dim; t.cand_V = d.V;
    }
    t.logit = logit;
    if (tail) {
      t.has_tail = 1;
      t.tail = TailParams{logit, tail->is_valid, tail->target, tail->feedback, tail->sample_u, t.all_u_t, tail->a_t,
                          tail->u_next, tail->action_score, tail->ce, B, A, d.E, nullptr};
      if (bpk_next) { t.tail.upk = bpk_next; t.tail.upk_NB = gpl.NB; }
    }
    if (merged) return launch_step_fused(f, t, st, ws.fz, ws.fz_bytes, ws.tsync, 256);
    return launch_text_score_fused(t, st, ws.tsync, 256);
  }
  if (ctx_k) {
    // model.py:395-396 with the per-episode projections of ctx (sfb_follower_project_ctx): the text attention reads
    // h_1 directly (scores = (ctx W_in) . h, output = sum alpha (ctx W_out_c^T) = W_out_c wc), so NO projection sits
    // between the LSTM cell and the attention.  hh = W_out_h h is its own small projection that triggers its
    // dependents only after its dependency wait; the attention (which needs nothing from it) then runs CONCURRENTLY
    // with it and waits for it only before its epilogue, where h~ = tanh(W_out_c wc + hh) is formed.  The g projection
    // also computes the next step's query from h_1 (extra tiles fed from a second activation source).
    const bool q_with_hh = g_q_with_hh != 0;   // where the next step's query is computed: with hh (off the chain) or with g
    {
      PkParams q{};
      q.a_pk = base + P.a_th + pk_weight_bytes(d.H, P.nkb_h); q.nkb = kblocks(d.H);   // rows [W_out[:, H:2H] ; M_q]
      q.g.nseg = 1;
      q.g.seg[0] = GemmSeg{ws.h1d, d.H, nullptr, nullptr, 0, nullptr, 0, d.H, 0};
      q.g.M = B; q.g.N = d.H; q.g.out = ws.th; q.g.ldo = d.H;
      if (q_next && q_with_hh) {   // tiles H/128.. = M_q rows fed with the un-dropped h_1
        q.g.N = d.H + d.F;
        q.g.n_split = d.H; q.g.out2 = q_next; q.g.ldo2 = d.F; q.g.bias2 = b_q;
        q.alt_tile0 = d.H / 128;
        q.alt_seg = GemmSeg{h1, d.H, nullptr, nullptr, 0, nullptr, 0, d.H, 0};
      }
      q.late_trigger = 1;
      SFB_PROPAGATE(launch_gemm_pk(q, st, ws.pk, ws.pk_bytes));
    }
    {
      AttnParams a{};
      a.q = ws.h1d; a.ldq = d.H; a.R = L; a.D = d.H;
      a.segA = ctx_o; a.strideA_b = (long long)L * d.H; a.strideA_r = d.H; a.lenA = d.H; a.lenB = 0;
      a.keyA = ctx_k; a.strideK_b = (long long)L * d.H; a.strideK_r = d.H;
      a.mask = ctx_mask; a.ldmask = L;
      a.out = ws.htilde; a.ldo = d.H; a.alpha = alpha; a.ldalpha = L;
      a.defer_wait = 1;
      a.post_add = ws.th; a.ld_post = d.H; a.post_tanh = 1;   // h~ = tanh(W_out_c wc + hh), once, in the epilogue
      SFB_PROPAGATE(launch_soft_dot_attention(a, B, ws.at, ws.at_bytes, st));
    }
    {
      PkParams q{};
      q.a_pk = base + P.a_g; q.b_pk = nullptr; q.nkb = kblocks(d.H);
      q.g.nseg = 1;
      q.g.seg[0] = GemmSeg{ws.htilde, d.H, nullptr, nullptr, 0, nullptr, 0, d.H, 0};
      q.g.M = B; q.g.out = ws.g; q.g.ldo = ws.ldg; q.g.bias0 = b_g;
      const int g_tiles = (d.E + 1 + 127) / 128;
      if (q_next && !q_with_hh) {   // tiles g_tiles.. = M_q rows, fed with the un-dropped h_1: the next step's visual query
        SFB_CHECK_ARG(P.a_gq == P.a_g + pk_weight_bytes(d.E + 1, P.nkb_h), "packed layout: g / q operand not contiguous");
        q.g.N = g_tiles * 128 + d.F;
        q.g.n_split = g_tiles * 128; q.g.n1_valid = d.E + 1; q.g.out2 = q_next; q.g.ldo2 = d.F; q.g.bias2 = b_q;
        q.alt_tile0 = g_tiles;
        q.alt_seg = GemmSeg{h1, d.H, nullptr, nullptr, 0, nullptr, 0, d.H, 0};
      } else {
        q.g.N = d.E + 1;
      }
      SFB_PROPAGATE(launch_gemm_pk(q, st, ws.pk, ws.pk_bytes));
    }
  } else {
    // model.py:395  text attention: [t | W_out_h h1d] in one projection — and, software-pipelined across the
    // recurrence, the NEXT step's visual query M_q h_1 + b_q in the same launch (eval mode: h1d == h_1)
    if (q_next && !drop_h) {
      SFB_PROPAGATE(proj(base + P.a_th, ws.h1d, d.H, d.H, 2 * d.H + d.F, ws.th, 2 * d.H, nullptr, nullptr, 0, 0, 2 * d.H, q_next,
                         d.F, b_q));
    } else {
      SFB_PROPAGATE(proj(base + P.a_th, ws.h1d, d.H, d.H, 2 * d.H, ws.th, 2 * d.H, nullptr, nullptr, 0, 0, 0, nullptr, 0, nullptr));
    }
    {
      AttnParams a{};
      a.q = ws.th; a.ldq = 2 * d.H; a.R = L; a.D = d.H;
      a.segA = ctx; a.strideA_b = (long long)L * d.H; a.strideA_r = d.H; a.lenA = d.H; a.lenB = 0;
      a.mask = ctx_mask; a.ldmask = L;
      a.out = ws.wc; a.ldo = d.H; a.alpha = alpha; a.ldalpha = L;
      SFB_PROPAGATE(launch_soft_dot_attention(a, B, ws.at, ws.at_bytes, st));
    }
    SFB_PROPAGATE(proj(base + P.a_wc, ws.wc, d.H, d.H, d.H, ws.htilde, d.H, nullptr, ws.th + d.H, 2 * d.H, 1, 0, nullptr, 0, nullptr));
    // model.py:396  logit = decoder2action(h_tilde, all_u_t):  g = M_g h~ + b_g (column E = the per-row constant)
    SFB_PROPAGATE(proj(base + P.a_g, ws.htilde, d.H, d.H, d.E + 1, ws.g, ws.ldg, b_g, nullptr, 0, 0, 0, nullptr, 0, nullptr));
  }
  ScoringParams sp{};
  sp.all_u_t = all_u_t; sp.g = ws.g; sp.tp = nullptr; sp.ldg = ws.ldg; sp.logit = logit; sp.B = B; sp.A = A; sp.E = d.E; sp.D = d.D;
  if (act_gather) {
    sp.all_u_t = nullptr; sp.cand_table = act->feat_table; sp.vp_idx = act->vp_idx; sp.cand_view = act->cand_view;
    sp.cand_trig = act->cand_trig; sp.img_dim = act->img_dim; sp.cand_V = d.V;
  }
  if (tail) {   // follower.py:476-505 fused behind the logits
    sp.has_tail = 1;
    sp.tail = TailParams{logit, tail->is_valid, tail->target, tail->feedback, tail->sample_u, sp.all_u_t, tail->a_t,
                         tail->u_next, tail->action_score, tail->ce, B, A, d.E, nullptr};
    if (bpk_next && gpl.nz == 1) {   // the chosen candidate row also in packed form: the u_prev blocks of the next step's gate GEMM
      sp.tail.upk = bpk_next; sp.tail.upk_NB = gpl.NB;
    }
  }
  SFB_PROPAGATE(launch_action_scoring(sp, st));
  if (q_next && drop_h && !ctx_k)   // train mode: the next query needs the un-dropped h_1 -> its own projection
    SFB_PROPAGATE(proj(base + P.a_q, h1, d.H, d.H, d.F, q_next, d.F, b_q, nullptr, 0, 0, 0, nullptr, 0, nullptr));
  return 0;
}

/* ---------------------------------------------------------------- follower decode step: backward */
namespace {
struct BwdWs {
  float *dg, *dsum, *th, *v, *r, *dth, *prod, *dht, *dz, *dwc_dh, *t, *dt, *wc, *dh1d, *dgates, *x, *dfeat, *dq, *tv, *dtv;
  size_t bytes;
};
BwdWs carve_bwd(const sfb_dims& d, int B, void* ws) {
  BwdWs w;
  Carver c(ws);
  w.dg = c.take((size_t)B * d.E); w.dsum = c.take(B);
  w.th = c.take((size_t)B * d.D); w.v = c.take((size_t)B * d.D); w.r = c.take((size_t)B * d.D);
  w.dth = c.take((size_t)B * d.D); w.prod = c.take((size_t)B * d.D);
  w.dht = c.take((size_t)B * d.H); w.dz = c.take((size_t)B * d.H); w.dwc_dh = c.take((size_t)B * 2 * d.H);
  w.t = c.take((size_t)B * d.H); w.dt = c.take((size_t)B * d.H); w.wc = c.take((size_t)B * d.H); w.dh1d = c.take((size_t)B * d.H);
  w.dgates = c.take((size_t)B * 4 * d.H); w.x = c.take((size_t)B * (d.E + d.F));
  w.dfeat = c.take((size_t)B * d.F); w.dq = c.take((size_t)B * d.F);
  w.tv = c.take((size_t)B * d.D); w.dtv = c.take((size_t)B * d.D);
  w.bytes = c.off;
  return w;
}
// out[M,N] = x[M,K] · W^T (kn = 0: W is [N,K]) or x · W (kn = 1: W is [K,N]), + bias, + padd
int32_t bgemm(int M, int N, int K, const float* x, int ldx, const float* W, int ldw, int kn, float* out, int ldo,
              const float* bias, const float* padd, int ld_padd, cudaStream_t st) {
  GemmParams g{};
  g.nseg = 1;
  g.seg[0] = GemmSeg{x, ldx, nullptr, nullptr, 0, W, ldw, K, kn};
  g.M = M; g.N = N; g.splitk = gemm_pick_splitk(M, N, K, device_num_sms());
  g.out = out; g.ldo = ldo; g.bias0 = bias; g.padd = padd; g.ld_padd = ld_padd;
  return launch_gemm(g, st);
}
int32_t outer(const float* Y, int ldy, const float* X, int ldx, int B, int N, int K, float* out, int ldo, int acc, cudaStream_t st) {
  if (!out) return 0;
  OuterParams o{Y, ldy, X, ldx, B, N, K, out, ldo, acc};
  return launch_outer_accum(o, st);
}
}  // namespace

/* ---------------------------------------------------------------- EncoderLSTM backward (BPTT over the taped forward) */
namespace {
struct EncBwdWs { float *dgates, *xemb, *dpre, *dh[2], *dc[2], *pass; size_t bytes; };
EncBwdWs carve_enc_bwd(int Hd, int Ew, int B, int maxlen, void* p) {
  EncBwdWs w;
  Carver c(p);
  w.dgates = c.take((size_t)maxlen * B * 4 * Hd);
  w.xemb = c.take((size_t)maxlen * B * Ew);
  w.dpre = c.take((size_t)B * Hd);
  for (int i = 0; i < 2; ++i) { w.dh[i] = c.take((size_t)B * Hd); w.dc[i] = c.take((size_t)B * Hd); }
  w.pass = c.take((size_t)B * Hd);
  w.bytes = c.off;
  return w;
}
}  // namespace

size_t sfb_encoder_lstm_bwd_workspace_bytes(int32_t Hd, int32_t Ew, int32_t B, int32_t maxlen) {
  if (Hd < 1 || Ew < 1 || B < 1 || maxlen < 1) return 0;
  return carve_enc_bwd(Hd, Ew, B, maxlen, nullptr).bytes;
}

int32_t sfb_encoder_lstm_bwd(const sfb_encoder_weights* w, int32_t Hd, int32_t Ew, int32_t B, int32_t maxlen,
                             const int32_t* seq, const int32_t* lengths, const float* drop_embed, const void* tape_mem,
                             const float* decoder_init, const float* g_ctx, const float* g_decoder_init, const float* g_c_t,
                             const sfb_encoder_grads* gr, int32_t accumulate, void* workspace, size_t workspace_bytes,
                             void* stream) {
  reset_launch_count();
  SFB_CHECK_ARG(w && seq && lengths && tape_mem && decoder_init && gr, "NULL argument");
  SFB_CHECK_ARG(Hd >= 4 && (Hd % 4) == 0 && Ew >= 4 && (Ew % 4) == 0 && B >= 1 && maxlen >= 1, "bad sizes");
  const EncTape tape = carve_enc_tape(Hd, B, maxlen, const_cast<void*>(tape_mem));
  const EncBwdWs bw = carve_enc_bwd(Hd, Ew, B, maxlen, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, bw.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int acc = accumulate ? 1 : 0, H = Hd;
  const size_t state = (size_t)B * Hd;
  const float* h_last = tape.h + (size_t)maxlen * state;
  // decoder_init = tanh(encoder2decoder(h_T))  (model.py:99)
  int cur = 0;
  if (g_decoder_init) {
    SFB_PROPAGATE(launch_tanh_bwd(g_decoder_init, decoder_init, bw.dpre, B * H, st));
    SFB_PROPAGATE(outer(bw.dpre, H, h_last, H, B, H, H, gr->e2d_w, H, acc, st));
    if (gr->e2d_b) SFB_PROPAGATE(launch_colsum(bw.dpre, H, B, H, gr->e2d_b, acc, st));
    SFB_PROPAGATE(bgemm(B, H, H, bw.dpre, H, w->e2d_w, H, 1, bw.dh[0], H, nullptr, nullptr, 0, st));   // dh_T = dpre W_e
  } else {
    SFB_CHECK_CUDA(cudaMemsetAsync(bw.dh[0], 0, state * sizeof(float), st));
  }
  if (g_c_t) SFB_CHECK_CUDA(cudaMemcpyAsync(bw.dc[0], g_c_t, state * sizeof(float), cudaMemcpyDeviceToDevice, st));
  else SFB_CHECK_CUDA(cudaMemsetAsync(bw.dc[0], 0, state * sizeof(float), st));
  // back-propagation through time (packed sequence: a row is active at step t iff t < lengths[row], model.py:89-90)
  for (int t = maxlen - 1; t >= 0; --t) {
    LstmSeqBwdParams p{};
    p.B = B; p.H = H; p.t = t; p.lengths = lengths;
    p.gates_act = tape.gates + (size_t)t * B * 4 * H; p.c_prev = tape.c + (size_t)t * state; p.c_cur = tape.c + (size_t)(t + 1) * state;
    p.dh_in = bw.dh[cur]; p.dc_in = bw.dc[cur];
    p.g_out = g_ctx ? g_ctx + (size_t)t * H : nullptr; p.ld_g_out = (long long)maxlen * H;
    p.dgates = bw.dgates + (size_t)t * B * 4 * H; p.dc_prev = bw.dc[cur ^ 1]; p.dh_pass = bw.pass;
    SFB_PROPAGATE(launch_lstm_seq_bwd(p, st));
    SFB_PROPAGATE(bgemm(B, H, 4 * H, p.dgates, 4 * H, w->w_hh[0], H, 1, bw.dh[cur ^ 1], H, nullptr, bw.pass, H, st));   // dh_{t-1} = dgates W_hh (+ pass-through)
    cur ^= 1;
  }
  // weight gradients, batched over all time steps: dW_hh = sum_t dgates_t^T h_{t-1}, dW_ih = sum_t dgates_t^T x_t
  const int rows = maxlen * B;
  SFB_PROPAGATE(outer(bw.dgates, 4 * H, tape.h, H, rows, 4 * H, H, gr->w_hh, H, acc, st));
  if (gr->w_ih) {
    SFB_PROPAGATE(launch_gather_embed(w->embedding, Ew, seq, drop_embed, bw.xemb, B, maxlen, st));
    SFB_PROPAGATE(outer(bw.dgates, 4 * H, bw.xemb, Ew, rows, 4 * H, Ew, gr->w_ih, Ew, acc, st));
  }
  if (gr->b_ih) SFB_PROPAGATE(launch_colsum(bw.dgates, 4 * H, rows, 4 * H, gr->b_ih, acc, st));
  if (gr->b_hh) SFB_PROPAGATE(launch_colsum(bw.dgates, 4 * H, rows, 4 * H, gr->b_hh, acc, st));
  return 0;
}

/* LSTMCell + VisualSoftDotAttention backward (model.py:389-394 / 431-434): shared by the follower step and the speaker
 * encoder step.  dh1d: gradient w.r.t. the DROPPED h_1 (NULL when nothing consumes it), feat / gates_act / alpha_v: what
 * the forward left behind. */
static int32_t vis_lstm_bwd(const sfb_dims& d, const sfb_vis_lstm_weights* wl, int B, const float* u_prev, const sfb_visual_source* vis,
                            const float* h0, const float* c0, const float* drop_x, const float* drop_h, const float* c1,
                            const float* alpha_v, const float* feat, const float* gates_act, const float* g_h1, const float* g_c1,
                            const float* dh1d, float* d_h0, float* d_c0, const sfb_follower_grads* gr, int acc, const BwdWs& w,
                            cudaStream_t st) {
  const int E = d.E, F = d.F, H = d.H, D = d.D;
  // ---- LSTMCell backward (model.py:393-394)
  {
    LstmBwdParams lp{B, H, gates_act, c0, c1, g_h1, g_c1, dh1d, drop_h, w.dgates, d_c0};
    SFB_PROPAGATE(launch_lstm_cell_bwd(lp, st));
  }
  if (gr->lstm_w_ih) {
    SFB_PROPAGATE(launch_assemble_x(u_prev, feat, drop_x, w.x, B, E, F, st));
    SFB_PROPAGATE(outer(w.dgates, 4 * H, w.x, E + F, B, 4 * H, E + F, gr->lstm_w_ih, E + F, acc, st));
  }
  SFB_PROPAGATE(outer(w.dgates, 4 * H, h0, H, B, 4 * H, H, gr->lstm_w_hh, H, acc, st));
  if (gr->lstm_b_ih) SFB_PROPAGATE(launch_colsum(w.dgates, 4 * H, B, 4 * H, gr->lstm_b_ih, acc, st));
  if (gr->lstm_b_hh) SFB_PROPAGATE(launch_colsum(w.dgates, 4 * H, B, 4 * H, gr->lstm_b_hh, acc, st));
  SFB_PROPAGATE(bgemm(B, F, 4 * H, w.dgates, 4 * H, wl->lstm_w_ih + E, E + F, 1, w.dfeat, F, nullptr, nullptr, 0, st));   // d(x_f) = dgates W_ih[:, E:]
  SFB_PROPAGATE(bgemm(B, H, 4 * H, w.dgates, 4 * H, wl->lstm_w_hh, H, 1, d_h0, H, nullptr, nullptr, 0, st));              // dh0 = dgates W_hh
  // ---- VisualSoftDotAttention backward (model.py:310-326)
  {
    AttnBwdParams a{};
    if (vis->visual) {
      a.segA = vis->visual; a.strideA_b = (long long)d.V * F; a.strideA_r = F; a.lenA = F; a.lenB = 0;
    } else {
      const int loc = F - vis->img_dim;
      a.segA = vis->feat_table; a.strideA_b = (long long)d.V * vis->img_dim; a.strideA_r = vis->img_dim; a.lenA = vis->img_dim; a.idxA = vis->vp_idx;
      a.segB = vis->loc_table; a.strideB_b = (long long)d.V * loc; a.strideB_r = loc; a.lenB = loc; a.idxB = vis->view_idx;
    }
    a.R = d.V; a.D = F; a.alpha = alpha_v; a.ldalpha = d.V;
    a.dout = w.dfeat; a.lddout = F; a.dout_scale = drop_x ? drop_x + E : nullptr; a.ldscale = E + F;
    a.dq = w.dq; a.lddq = F;
    SFB_PROPAGATE(launch_attn_bwd(a, B, st));
  }
  SFB_PROPAGATE(bgemm(B, D, H, h0, H, wl->va_w_h, H, 0, w.tv, D, wl->va_b_h, nullptr, 0, st));               // t_v = W_h h0 + b_h
  SFB_PROPAGATE(outer(w.tv, D, w.dq, F, B, D, F, gr->va_w_v, F, acc, st));                                   // dW_v = t_v^T dq
  SFB_PROPAGATE(bgemm(B, D, F, w.dq, F, wl->va_w_v, F, 0, w.dtv, D, nullptr, nullptr, 0, st));              // dt_v = W_v dq
  SFB_PROPAGATE(outer(w.dtv, D, h0, H, B, D, H, gr->va_w_h, H, acc, st));
  if (gr->va_b_h) SFB_PROPAGATE(launch_colsum(w.dtv, D, B, D, gr->va_b_h, acc, st));
  SFB_PROPAGATE(bgemm(B, H, D, w.dtv, D, wl->va_w_h, H, 1, d_h0, H, nullptr, d_h0, H, st));                  // dh0 += dt_v W_h
  return 0;
}

/* ---------------------------------------------------------------- follower decode step: backward (entry points) */
size_t sfb_follower_step_bwd_workspace_bytes(const sfb_dims* dims, int32_t B, int32_t L, int32_t A) {
  (void)L; (void)A;
  if (!dims || B < 1) return 0;
  return carve_bwd(*dims, B, nullptr).bytes;
}

int32_t sfb_follower_step_bwd(const sfb_dims* dims, const sfb_vis_lstm_weights* wl, const sfb_softdot_weights* wt,
                              const sfb_scoring_weights* wsc, int32_t B, int32_t L, int32_t A, const float* u_prev,
                              const sfb_action_source* act, const sfb_visual_source* vis, const float* h0, const float* c0,
                              const float* ctx, const uint8_t* ctx_mask, const float* drop_x, const float* drop_h,
                              const float* c1, const float* alpha, const float* alpha_v, const void* fwd_workspace,
                              const float* g_h1, const float* g_c1, const float* g_logit, float* d_h0, float* d_c0,
                              float* d_ctx, const sfb_follower_grads* gr, int32_t accumulate, void* workspace,
                              size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_CHECK_ARG(wl && wt && wsc && act && vis && gr, "NULL struct argument");
  SFB_CHECK_ARG(u_prev && h0 && c0 && ctx && c1 && alpha && alpha_v && fwd_workspace && d_h0 && d_c0 && d_ctx, "NULL tensor argument");
  SFB_CHECK_ARG(B >= 1 && L >= 1 && A >= 1, "B, L, A >= 1");
  const sfb_dims& d = *dims;
  const FollowerWs fw = carve_follower(d, B, L, A, const_cast<void*>(fwd_workspace));   // where the forward left its intermediates
  const BwdWs w = carve_bwd(d, B, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, w.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int acc = accumulate ? 1 : 0;
  const float* h_tilde = fw.htilde; const float* h1d = fw.h1d; const float* feat = fw.feat; const float* gates_act = fw.gates_act;
  const int E = d.E, F = d.F, H = d.H, D = d.D;

  // ---- EltwiseProdScoring backward (model.py:342-352)
  if (g_logit) {
    ScoreBwdParams sp{};
    sp.dlogit = g_logit; sp.B = B; sp.A = A; sp.E = E; sp.dg = w.dg; sp.dsum = w.dsum;
    if (act->all_u_t) sp.all_u_t = act->all_u_t;
    else { sp.cand_table = act->feat_table; sp.vp_idx = act->vp_idx; sp.cand_view = act->cand_view; sp.cand_trig = act->cand_trig;
           sp.img_dim = act->img_dim; sp.cand_V = d.V; }
    SFB_PROPAGATE(launch_score_bwd(sp, st));
    SFB_PROPAGATE(bgemm(B, D, H, h_tilde, H, wsc->w_h, H, 0, w.th, D, wsc->b_h, nullptr, 0, st));          // th = W_h h~ + b_h
    SFB_PROPAGATE(bgemm(B, D, E, w.dg, E, wsc->w_a, E, 0, w.v, D, nullptr, nullptr, 0, st));               // v = W_a dg
    SFB_PROPAGATE(launch_scoring_mid(w.th, w.v, wsc->w_out, wsc->b_a, w.dsum, w.r, w.dth, w.prod, B, D, st));
    SFB_PROPAGATE(outer(w.r, D, w.dg, E, B, D, E, gr->sc_w_a, E, acc, st));                                 // dW_a = r^T dg
    SFB_PROPAGATE(outer(w.r, D, w.dsum, 1, B, D, 1, gr->sc_b_a, 1, acc, st));                               // db_a = r^T dsum
    if (gr->sc_w_out) SFB_PROPAGATE(launch_colsum(w.prod, D, B, D, gr->sc_w_out, acc, st));                 // dw_o
    if (gr->sc_b_out) SFB_PROPAGATE(launch_colsum(w.dsum, 1, B, 1, gr->sc_b_out, acc, st));                 // db_o
    SFB_PROPAGATE(outer(w.dth, D, h_tilde, H, B, D, H, gr->sc_w_h, H, acc, st));                            // dW_h' = dth^T h~
    if (gr->sc_b_h) SFB_PROPAGATE(launch_colsum(w.dth, D, B, D, gr->sc_b_h, acc, st));
    SFB_PROPAGATE(bgemm(B, H, D, w.dth, D, wsc->w_h, H, 1, w.dht, H, nullptr, nullptr, 0, st));            // dh~ = dth W_h'
  } else {
    SFB_CHECK_CUDA(cudaMemsetAsync(w.dht, 0, (size_t)B * H * sizeof(float), st));
  }
  // ---- SoftDotAttention backward (model.py:122-143)
  SFB_PROPAGATE(launch_tanh_bwd(w.dht, h_tilde, w.dz, B * H, st));
  SFB_PROPAGATE(bgemm(B, 2 * H, H, w.dz, H, wt->w_out, 2 * H, 1, w.dwc_dh, 2 * H, nullptr, nullptr, 0, st));   // [dwc | dh1d] = dz W_out
  SFB_PROPAGATE(bgemm(B, H, H, h1d, H, wt->w_in, H, 0, w.t, H, nullptr, nullptr, 0, st));                      // t = W_in h1d
  {
    AttnBwdParams a{};
    a.segA = ctx; a.strideA_b = (long long)L * H; a.strideA_r = H; a.lenA = H; a.lenB = 0;
    a.mask = ctx_mask; a.ldmask = L; a.R = L; a.D = H;
    a.alpha = alpha; a.ldalpha = L; a.dout = w.dwc_dh; a.lddout = 2 * H; a.qv = w.t; a.ldq = H;
    a.dq = w.dt; a.lddq = H; a.drows = d_ctx; a.wsum = w.wc; a.ldwsum = H;
    SFB_PROPAGATE(launch_attn_bwd(a, B, st));
  }
  SFB_PROPAGATE(outer(w.dt, H, h1d, H, B, H, H, gr->w_in, H, acc, st));                                      // dW_in = dt^T h1d
  SFB_PROPAGATE(outer(w.dz, H, w.wc, H, B, H, H, gr->w_out, 2 * H, acc, st));                                 // dW_out[:, :H] = dz^T wc
  SFB_PROPAGATE(outer(w.dz, H, h1d, H, B, H, H, gr->w_out ? gr->w_out + H : nullptr, 2 * H, acc, st));       // dW_out[:, H:] = dz^T h1d
  SFB_PROPAGATE(bgemm(B, H, H, w.dt, H, wt->w_in, H, 1, w.dh1d, H, nullptr, w.dwc_dh + H, 2 * H, st));       // dh1d = dt W_in + dz W_out_h
  return vis_lstm_bwd(d, wl, B, u_prev, vis, h0, c0, drop_x, drop_h, c1, alpha_v, feat, gates_act, g_h1, g_c1, w.dh1d, d_h0, d_c0,
                      gr, acc, w, st);
}

int32_t sfb_sf_search_update(const sfb_sf_search_state* st, const sfb_nav_tables* nav, int32_t B, int32_t iter,
                             int32_t episode_len, int32_t completion_size, const float* lp, void* stream) {
  reset_launch_count();
  SFB_CHECK_ARG(st && nav && lp, "NULL argument");
  SFB_CHECK_ARG(st->beam_node && st->c_score && st->c_node && st->c_exp && st->h_score && st->h_node && st->h_exp && st->d_score &&
                    st->d_node && st->n_done && st->n_nodes && st->node_parent && st->node_state && st->node_action &&
                    st->node_count && st->node_slot && st->node_score && st->trav && st->flags, "search state: NULL array");
  SFB_CHECK_ARG(nav->next && nav->nvalid && nav->S >= 1 && nav->A >= 1, "navigation tables missing");
  SfSearchParams p{};
  p.B = B; p.A = nav->A; p.S = nav->S; p.M = st->max_nodes; p.max_iter = st->max_iter; p.episode_len = episode_len;
  p.completion_size = completion_size; p.iter = iter; p.lp = lp; p.nav_next = nav->next; p.nav_nvalid = nav->nvalid;
  p.beam_node = st->beam_node; p.c_score = st->c_score; p.c_node = st->c_node; p.c_exp = st->c_exp;
  p.h_score = st->h_score; p.h_node = st->h_node; p.h_exp = st->h_exp; p.d_score = st->d_score; p.d_node = st->d_node;
  p.n_done = st->n_done; p.n_nodes = st->n_nodes; p.node_parent = st->node_parent; p.node_state = st->node_state;
  p.node_action = st->node_action; p.node_count = st->node_count; p.node_slot = st->node_slot; p.node_score = st->node_score;
  p.trav = st->trav; p.flags = st->flags;
  return launch_sf_search_update(p, static_cast<cudaStream_t>(stream));
}

/* ---------------------------------------------------------------- speaker modules: backward (train_speaker.py, speaker.py:376-395) */
size_t sfb_speaker_encoder_step_bwd_workspace_bytes(const sfb_dims* dims, int32_t B) {
  if (!dims || B < 1) return 0;
  return carve_bwd(*dims, B, nullptr).bytes;
}

int32_t sfb_speaker_encoder_step_bwd(const sfb_dims* dims, const sfb_vis_lstm_weights* wl, int32_t B, const float* action_embedding,
                                     const sfb_visual_source* vis, const float* h0, const float* c0, const float* drop_x,
                                     const float* c1, const void* fwd_workspace, const float* g_h1, const float* g_c1,
                                     float* d_h0, float* d_c0, const sfb_follower_grads* gr, int32_t accumulate,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_CHECK_ARG(wl && vis && gr, "NULL struct argument");
  SFB_CHECK_ARG(action_embedding && h0 && c0 && c1 && fwd_workspace && d_h0 && d_c0, "NULL tensor argument");
  SFB_CHECK_ARG(B >= 1 && dims->V <= dims->D, "B >= 1, V <= D (the forward keeps the attention weights in a [B,D] region)");
  const sfb_dims& d = *dims;
  const FollowerWs fw = carve_follower(d, B, 1, 1, const_cast<void*>(fwd_workspace));   // where the forward left feat / gates / alpha_v
  const BwdWs w = carve_bwd(d, B, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, w.bytes));
  return vis_lstm_bwd(d, wl, B, action_embedding, vis, h0, c0, drop_x, nullptr, c1, fw.tp, fw.feat, fw.gates_act, g_h1, g_c1, nullptr,
                      d_h0, d_c0, gr, accumulate ? 1 : 0, w, static_cast<cudaStream_t>(stream));
}

namespace {
struct SpkDecBwdWs { float *dht, *dz, *dwc_dh, *t, *dt, *wc, *dh1d, *dgates, *x, *gT; size_t bytes; };
SpkDecBwdWs carve_spkdec_bwd(int H, int Ew, int vocab, int B, void* p) {
  SpkDecBwdWs w;
  Carver c(p);
  w.dht = c.take((size_t)B * H); w.dz = c.take((size_t)B * H); w.dwc_dh = c.take((size_t)B * 2 * H);
  w.t = c.take((size_t)B * H); w.dt = c.take((size_t)B * H); w.wc = c.take((size_t)B * H); w.dh1d = c.take((size_t)B * H);
  w.dgates = c.take((size_t)B * 4 * H); w.x = c.take((size_t)B * Ew);
  w.gT = c.take((size_t)vocab * B);
  w.bytes = c.off;
  return w;
}
}  // namespace

size_t sfb_speaker_decoder_step_bwd_workspace_bytes(int32_t H, int32_t Ew, int32_t vocab, int32_t B) {
  if (H < 1 || Ew < 1 || vocab < 1 || B < 1) return 0;
  return carve_spkdec_bwd(H, Ew, vocab, B, nullptr).bytes;
}

int32_t sfb_speaker_decoder_step_bwd(const sfb_speaker_decoder_weights* w, int32_t H, int32_t Ew, int32_t vocab, int32_t B,
                                     int32_t T, const int32_t* prev_word, const float* h0, const float* c0, const float* ctx,
                                     const uint8_t* ctx_mask, const float* drop_e, const float* drop_h, const float* c1,
                                     const float* alpha, const void* fwd_workspace, const float* g_h1, const float* g_c1,
                                     const float* g_logit, float* d_h0, float* d_c0, float* d_ctx,
                                     const sfb_speaker_decoder_grads* gr, int32_t accumulate, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_CHECK_ARG(w && gr && prev_word && h0 && c0 && ctx && c1 && alpha && fwd_workspace && d_h0 && d_c0 && d_ctx, "NULL argument");
  SFB_CHECK_ARG(H >= 4 && (H % 4) == 0 && Ew >= 4 && (Ew % 4) == 0 && vocab >= 1 && B >= 1 && T >= 1, "bad sizes");
  const SpkDecWs fw = carve_spkdec(H, Ew, B, T, const_cast<void*>(fwd_workspace));
  const SpkDecBwdWs bw = carve_spkdec_bwd(H, Ew, vocab, B, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, bw.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int acc = accumulate ? 1 : 0;
  const float* h_tilde = fw.htilde; const float* h1d = fw.h1d;
  // ---- decoder2action backward (model.py:518): logit = W_voc h~ + b
  if (g_logit) {
    SFB_PROPAGATE(outer(g_logit, vocab, h_tilde, H, B, vocab, H, gr->w_voc, H, acc, st));
    if (gr->b_voc) SFB_PROPAGATE(launch_colsum(g_logit, vocab, B, vocab, gr->b_voc, acc, st));
    // dh~ = dlogit W_voc: a reduction over the vocabulary (991: not a multiple of 4, so not a skinny-GEMM K) — the
    // exact-fp32 outer-product kernel with the vocabulary as its reduction dimension
    SFB_PROPAGATE(launch_transpose(g_logit, vocab, B, vocab, bw.gT, B, st));
    SFB_PROPAGATE(outer(bw.gT, B, w->w_voc, H, vocab, B, H, bw.dht, H, 0, st));
  } else {
    SFB_CHECK_CUDA(cudaMemsetAsync(bw.dht, 0, (size_t)B * H * sizeof(float), st));
  }
  // ---- SoftDotAttention backward (model.py:122-143), as in the follower step
  SFB_PROPAGATE(launch_tanh_bwd(bw.dht, h_tilde, bw.dz, B * H, st));
  SFB_PROPAGATE(bgemm(B, 2 * H, H, bw.dz, H, w->attn.w_out, 2 * H, 1, bw.dwc_dh, 2 * H, nullptr, nullptr, 0, st));   // [dwc | dh1d] = dz W_out
  SFB_PROPAGATE(bgemm(B, H, H, h1d, H, w->attn.w_in, H, 0, bw.t, H, nullptr, nullptr, 0, st));                      // t = W_in h1d
  {
    AttnBwdParams a{};
    a.segA = ctx; a.strideA_b = (long long)T * H; a.strideA_r = H; a.lenA = H; a.lenB = 0;
    a.mask = ctx_mask; a.ldmask = T; a.R = T; a.D = H;
    a.alpha = alpha; a.ldalpha = T; a.dout = bw.dwc_dh; a.lddout = 2 * H; a.qv = bw.t; a.ldq = H;
    a.dq = bw.dt; a.lddq = H; a.drows = d_ctx; a.wsum = bw.wc; a.ldwsum = H;
    SFB_PROPAGATE(launch_attn_bwd(a, B, st));
  }
  SFB_PROPAGATE(outer(bw.dt, H, h1d, H, B, H, H, gr->w_in, H, acc, st));
  SFB_PROPAGATE(outer(bw.dz, H, bw.wc, H, B, H, H, gr->w_out, 2 * H, acc, st));
  SFB_PROPAGATE(outer(bw.dz, H, h1d, H, B, H, H, gr->w_out ? gr->w_out + H : nullptr, 2 * H, acc, st));
  SFB_PROPAGATE(bgemm(B, H, H, bw.dt, H, w->attn.w_in, H, 1, bw.dh1d, H, nullptr, bw.dwc_dh + H, 2 * H, st));       // dh1d = dt W_in + dz W_out_h
  // ---- LSTMCell backward (model.py:515); x = embedding(previous_word) (.) drop_e, the embedding is frozen GloVe
  {
    LstmBwdParams lp{B, H, fw.gates_act, c0, c1, g_h1, g_c1, bw.dh1d, drop_h, bw.dgates, d_c0};
    SFB_PROPAGATE(launch_lstm_cell_bwd(lp, st));
  }
  if (gr->lstm_w_ih) {
    SFB_PROPAGATE(launch_gather_embed(w->embedding, Ew, prev_word, drop_e, bw.x, B, 1, st));
    SFB_PROPAGATE(outer(bw.dgates, 4 * H, bw.x, Ew, B, 4 * H, Ew, gr->lstm_w_ih, Ew, acc, st));
  }
  SFB_PROPAGATE(outer(bw.dgates, 4 * H, h0, H, B, 4 * H, H, gr->lstm_w_hh, H, acc, st));
  if (gr->lstm_b_ih) SFB_PROPAGATE(launch_colsum(bw.dgates, 4 * H, B, 4 * H, gr->lstm_b_ih, acc, st));
  if (gr->lstm_b_hh) SFB_PROPAGATE(launch_colsum(bw.dgates, 4 * H, B, 4 * H, gr->lstm_b_hh, acc, st));
  return bgemm(B, H, 4 * H, bw.dgates, 4 * H, w->lstm_w_hh, H, 1, d_h0, H, nullptr, nullptr, 0, st);                // dh0 = dgates W_hh
}

/* The first half of a decode step alone (model.py:389-393): attention gather + gate GEMM + LSTM cell from carried
 * state, ONE launch of vis_lstm_fused_kernel.  carry_in holds the visual query and the packed [u_prev | . | h_0] blocks. */
int32_t sfb_follower_gather_lstm_fwd(const sfb_dims* dims, const sfb_vis_lstm_weights* wl, const void* packed, size_t packed_bytes,
                                     int32_t B, void* carry_in, const sfb_visual_source* vis, const float* c0, float* h1, float* c1,
                                     float* feature, float* alpha_v, void* workspace, size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_PROPAGATE(check_packable(*dims));
  SFB_CHECK_ARG(wl && packed && carry_in && vis && c0 && h1 && c1, "NULL argument");
  const sfb_dims& d = *dims;
  const FollowerPk P = layout_follower_pk(d);
  SFB_CHECK_ARG(packed_bytes >= P.bytes && (reinterpret_cast<uintptr_t>(packed) & 255u) == 0, "packed buffer too small / misaligned");
  FollowerWs ws = carve_follower(d, B, 1, 1, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, ws.bytes));
  const CarryLayout CL = layout_carry(d, B);
  const PkPlan gpl = gemm_pk_plan(B, 4 * d.H, P.nkb_gates, true, device_num_sms());
  const int lenA = vis->visual ? d.F : vis->img_dim, lenB = vis->visual ? 0 : d.F - vis->img_dim;
  const FusedPlan fpl = vis_lstm_fused_plan(B, d.H, P.nkb_gates, d.V, d.F, lenA, lenB, device_num_sms());
  SFB_CHECK_ARG(fpl.ok && gpl.nz == 1 && gpl.NB == fpl.NB, "gather + LSTM launch: unsupported shape (B <= 128, F = 2176)");
  const unsigned char* base = static_cast<const unsigned char*>(packed);
  FusedVisLstmParams f{};
  f.q = reinterpret_cast<const float*>(static_cast<char*>(carry_in) + CL.q); f.ldq = d.F; f.R = d.V; f.D = d.F;
  if (vis->visual) {
    f.segA = vis->visual; f.strideA_b = (long long)d.V * d.F; f.lenA = d.F; f.lenB = 0;
  } else {
    SFB_CHECK_ARG(vis->feat_table && vis->loc_table && vis->vp_idx && vis->view_idx, "gather visual source needs tables + indices");
    f.segA = vis->feat_table; f.strideA_b = (long long)d.V * vis->img_dim; f.lenA = vis->img_dim; f.idxA = vis->vp_idx;
    f.segB = vis->loc_table; f.strideB_b = (long long)d.V * lenB; f.lenB = lenB; f.idxB = vis->view_idx;
  }
  f.feat = feature ? feature : ws.feat; f.ldfeat = d.F; f.alpha = alpha_v; f.ldalpha = d.V;
  f.a_pk = base + P.a_gates; f.b_pk = reinterpret_cast<unsigned char*>(carry_in) + CL.bpk; f.nkb = P.nkb_gates;
  f.post_kb0 = kblocks(d.E); f.post_kb1 = kblocks(d.E) + kblocks(d.F); f.feat_kb0 = kblocks(d.E);
  LstmEpilogue& e = f.g.lstm;
  e.H = d.H; e.b_ih = wl->lstm_b_ih; e.b_hh = wl->lstm_b_hh; e.c0 = c0; e.h1 = h1; e.c1 = c1; e.gates_act = ws.gates_act;
  e.hpk_NB = gpl.NB; e.hpk_rows_per_z = gpl.rows_per_z;
  f.g.M = B; f.g.N = 4 * d.H; f.B = B;
  f.idx_dependent = vis->idx_dependent;
  f.pre_weight_free = g_fused_pre_weight;
  return launch_vis_lstm_fused(f, static_cast<cudaStream_t>(stream), ws.fz, ws.fz_bytes);
}

/* ---------------------------------------------------------------- speaker modules on the packed tcgen05 path */
namespace {
struct VisLstmPk { size_t a_q, b_q, a_gates, mq, bytes; int nkb_h, nkb_gates; };
VisLstmPk layout_vislstm_pk(const sfb_dims& d) {
  VisLstmPk L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t r = off; off += (bytes + 255) & ~size_t(255); return r; };
  L.nkb_h = kblocks(d.H);
  L.nkb_gates = kblocks(d.E) + kblocks(d.F) + kblocks(d.H);
  L.a_q = take(pk_weight_bytes(d.F, L.nkb_h));
  L.b_q = take((size_t)d.F * 4);
  L.a_gates = take(pk_weight_bytes(4 * d.H, L.nkb_gates));
  L.mq = take((size_t)d.F * d.H * 4);
  L.bytes = off;
  return L;
}
struct SpkDecPk { size_t a_gates, a_th, a_wc, a_voc, a_hh, tdec, bytes; int nkb_h, nkb_gates; };
SpkDecPk layout_spkdec_pk(int H, int Ew, int vocab) {
  SpkDecPk L{};
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t r = off; off += (bytes + 255) & ~size_t(255); return r; };
  L.nkb_h = kblocks(H);
  L.nkb_gates = kblocks(Ew) + kblocks(H);
  L.a_gates = take(pk_weight_bytes(4 * H, L.nkb_gates));
  L.a_th = take(pk_weight_bytes(2 * H, L.nkb_h));
  L.a_wc = take(pk_weight_bytes(H, L.nkb_h));
  L.a_voc = take(pk_weight_bytes(vocab, L.nkb_h));
  // hoisted input projection (speaker.py:158-182 feeds words from a fixed vocabulary through a frozen embedding): the
  // per-token table T = Emb W_ih^T [vocab, 4H] and the recurrent weights alone as a packed operand
  L.a_hh = take(pk_weight_bytes(4 * H, L.nkb_h));
  L.tdec = take((size_t)vocab * 4 * H * sizeof(float));
  L.bytes = off;
  return L;
}
int32_t pack_plain(const float* w, int ldw, int rows, int k, unsigned char* out, cudaStream_t st) {
  PackParams p{};
  p.nseg = 1;
  p.seg[0] = PackSeg{w, ldw, k, nullptr, 0, nullptr};
  p.ntile = (rows + 127) / 128; p.R = 128; p.rows_per_tile = 128; p.rows_valid = rows; p.lstm_H = 0;
  p.out = out;
  return launch_pack_rows(p, st);
}
}  // namespace

size_t sfb_vis_lstm_packed_bytes(const sfb_dims* dims) {
  if (!dims || check_dims(dims) != 0 || check_packable(*dims) != 0) return 0;
  return layout_vislstm_pk(*dims).bytes;
}

int32_t sfb_vis_lstm_pack_weights(const sfb_dims* dims, const sfb_vis_lstm_weights* wl, void* packed, size_t packed_bytes,
                                  void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_PROPAGATE(check_packable(*dims));
  SFB_CHECK_ARG(wl && packed, "NULL argument");
  const sfb_dims& d = *dims;
  const VisLstmPk L = layout_vislstm_pk(d);
  SFB_CHECK_ARG(packed_bytes >= L.bytes && (reinterpret_cast<uintptr_t>(packed) & 255u) == 0, "packed buffer too small / misaligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* base = static_cast<unsigned char*>(packed);
  float* mq = reinterpret_cast<float*>(base + L.mq);
  SFB_PROPAGATE(launch_fold(wl->va_w_v, d.F, nullptr, wl->va_w_h, d.H, wl->va_b_h, d.D, d.F, d.H, mq, d.H,
                            reinterpret_cast<float*>(base + L.b_q), nullptr, st));
  SFB_PROPAGATE(pack_plain(mq, d.H, d.F, d.H, base + L.a_q, st));
  PackParams p{};
  p.nseg = 3;
  p.seg[0] = PackSeg{wl->lstm_w_ih, d.E + d.F, d.E, nullptr, 0, nullptr};
  p.seg[1] = PackSeg{wl->lstm_w_ih + d.E, d.E + d.F, d.F, nullptr, 0, nullptr};
  p.seg[2] = PackSeg{wl->lstm_w_hh, d.H, d.H, nullptr, 0, nullptr};
  p.ntile = d.H / 32; p.R = 128; p.rows_per_tile = 128; p.rows_valid = 4 * d.H; p.lstm_H = d.H;
  p.out = base + L.a_gates;
  return launch_pack_rows(p, st);
}

int32_t sfb_speaker_encoder_step_packed_fwd(const sfb_dims* dims, const sfb_vis_lstm_weights* w, const void* packed,
                                            size_t packed_bytes, int32_t B, const float* action_embedding,
                                            const sfb_visual_source* vis, const float* h0, const float* c0,
                                            const float* drop_x, float* h1, float* c1, void* workspace,
                                            size_t workspace_bytes, void* stream) {
  reset_launch_count();
  SFB_PROPAGATE(check_dims(dims));
  SFB_PROPAGATE(check_packable(*dims));
  SFB_CHECK_ARG(w && vis && packed && action_embedding && h0 && c0 && h1 && c1, "NULL argument");
  SFB_CHECK_ARG(B >= 1, "B >= 1");
  const sfb_dims& d = *dims;
  const VisLstmPk P = layout_vislstm_pk(d);
  SFB_CHECK_ARG(packed_bytes >= P.bytes && (reinterpret_cast<uintptr_t>(packed) & 255u) == 0, "packed buffer too small / misaligned");
  FollowerWs ws = carve_follower(d, B, 1, 1, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, ws.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned char* base = static_cast<const unsigned char*>(packed);
  // model.py:431  feature = visual_attention_layer(h, world_state): q = M_q h + b_q on tcgen05
  {
    PkParams q{};
    q.a_pk = base + P.a_q; q.nkb = P.nkb_h;
    q.g.nseg = 1;
    q.g.seg[0] = GemmSeg{h0, d.H, nullptr, nullptr, 0, nullptr, 0, d.H, 0};
    q.g.M = B; q.g.N = d.F; q.g.out = ws.q; q.g.ldo = d.F; q.g.bias0 = reinterpret_cast<const float*>(base + P.b_q);
    SFB_PROPAGATE(launch_gemm_pk(q, st, ws.pk, ws.pk_bytes));
  }
  const PkPlan gpl = gemm_pk_plan(B, 4 * d.H, P.nkb_gates, true, device_num_sms());
  {
    AttnParams pk{};
    pk.pk_out = ws.bpk; pk.pk_kb0 = kblocks(d.E); pk.pk_nkb = P.nkb_gates; pk.pk_NB = gpl.NB; pk.pk_rows_per_z = gpl.rows_per_z;
    pk.pk_scale = drop_x ? drop_x + d.E : nullptr; pk.pk_ldscale = d.E + d.F;
    pk.has_side = 1;
    PackParams& side = pk.side;
    side.nseg = 3;
    side.seg[0] = PackSeg{action_embedding, d.E, d.E, drop_x, drop_x ? d.E + d.F : 0, nullptr};
    side.seg[1] = PackSeg{nullptr, d.F, d.F, nullptr, 0, nullptr};
    side.seg[2] = PackSeg{h0, d.H, d.H, nullptr, 0, nullptr};
    side.ntile = gpl.nz; side.R = gpl.NB; side.rows_per_tile = gpl.rows_per_z; side.rows_valid = B; side.lstm_H = 0;
    side.out = ws.bpk;
    SFB_PROPAGATE(visual_attend(d, B, ws.q, *vis, ws.feat, d.V <= d.D ? ws.tp : nullptr, ws.av, ws.av_bytes, st, &pk));
  }
  // model.py:432-434  LSTMCell(drop(cat(action_embedding, feature)), (h, c))
  PkParams q{};
  q.a_pk = base + P.a_gates; q.b_pk = ws.bpk; q.nkb = P.nkb_gates;
  q.g.M = B; q.g.N = 4 * d.H;
  LstmEpilogue& e = q.g.lstm;
  e.H = d.H; e.b_ih = w->lstm_b_ih; e.b_hh = w->lstm_b_hh; e.c0 = c0;
  e.h1 = h1; e.c1 = c1; e.gates_act = ws.gates_act;
  return launch_gemm_pk(q, st, ws.pk, ws.pk_bytes);
}

size_t sfb_speaker_decoder_packed_bytes(int32_t H, int32_t Ew, int32_t vocab) {
  if (H < 128 || (H % 128) != 0 || Ew < 4 || (Ew % 4) != 0 || vocab < 1 || vocab > 4096) return 0;
  return layout_spkdec_pk(H, Ew, vocab).bytes;
}

int32_t sfb_speaker_decoder_pack_weights(const sfb_speaker_decoder_weights* w, int32_t H, int32_t Ew, int32_t vocab,
                                         void* packed, size_t packed_bytes, void* stream) {
  reset_launch_count();
  SFB_CHECK_ARG(w && packed, "NULL argument");
  SFB_CHECK_ARG(sfb_speaker_decoder_packed_bytes(H, Ew, vocab) != 0, "packed path needs H % 128 == 0, Ew % 4 == 0, vocab <= 4096");
  const SpkDecPk L = layout_spkdec_pk(H, Ew, vocab);
  SFB_CHECK_ARG(packed_bytes >= L.bytes && (reinterpret_cast<uintptr_t>(packed) & 255u) == 0, "packed buffer too small / misaligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned char* base = static_cast<unsigned char*>(packed);
  PackParams p{};
  p.nseg = 2;
  p.seg[0] = PackSeg{w->lstm_w_ih, Ew, Ew, nullptr, 0, nullptr};
  p.seg[1] = PackSeg{w->lstm_w_hh, H, H, nullptr, 0, nullptr};
  p.ntile = H / 32; p.R = 128; p.rows_per_tile = 128; p.rows_valid = 4 * H; p.lstm_H = H;
  p.out = base + L.a_gates;
  SFB_PROPAGATE(launch_pack_rows(p, st));
  SFB_PROPAGATE(pack_plain(w->attn.w_in, H, H, H, base + L.a_th, st));
  SFB_PROPAGATE(pack_plain(w->attn.w_out + H, 2 * H, H, H, base + L.a_th + pk_weight_bytes(H, L.nkb_h), st));
  SFB_PROPAGATE(pack_plain(w->attn.w_out, 2 * H, H, H, base + L.a_wc, st));
  SFB_PROPAGATE(pack_plain(w->w_voc, H, vocab, H, base + L.a_voc, st));
  {
    PackParams ph{};
    ph.nseg = 1;
    ph.seg[0] = PackSeg{w->lstm_w_hh, H, H, nullptr, 0, nullptr};
    ph.ntile = H / 32; ph.R = 128; ph.rows_per_tile = 128; ph.rows_valid = 4 * H; ph.lstm_H = H;
    ph.out = base + L.a_hh;
    SFB_PROPAGATE(launch_pack_rows(ph, st));
  }
  // T[v, :] = W_ih emb[v]  (exact fp32 products; biases are added by the cell epilogue as before)
  return bgemm(vocab, 4 * H, Ew, w->embedding, Ew, w->lstm_w_ih, Ew, 0, reinterpret_cast<float*>(base + L.tdec), 4 * H, nullptr, nullptr,
               0, st);
}

int32_t sfb_speaker_decoder_step_packed_fwd(const sfb_speaker_decoder_weights* w, const void* packed, size_t packed_bytes,
                                            int32_t H, int32_t Ew, int32_t vocab, int32_t B, int32_t T,
                                            const int32_t* prev_word, const float* h0, const float* c0, const float* ctx,
                                            const uint8_t* ctx_mask, const float* drop_e, const float* drop_h, float* h1,
                                            float* c1, float* alpha, float* logit, void* workspace, size_t workspace_bytes,
                                            void* stream) {
  reset_launch_count();
  SFB_CHECK_ARG(w && packed && prev_word && h0 && c0 && ctx && h1 && c1 && logit, "NULL argument");
  SFB_CHECK_ARG(sfb_speaker_decoder_packed_bytes(H, Ew, vocab) != 0, "packed path needs H % 128 == 0, Ew % 4 == 0, vocab <= 4096");
  SFB_CHECK_ARG(B >= 1 && T >= 1, "B, T >= 1");
  const SpkDecPk P = layout_spkdec_pk(H, Ew, vocab);
  SFB_CHECK_ARG(packed_bytes >= P.bytes && (reinterpret_cast<uintptr_t>(packed) & 255u) == 0, "packed buffer too small / misaligned");
  SpkDecWs ws = carve_spkdec(H, Ew, B, T, workspace);
  SFB_PROPAGATE(check_ws(workspace, workspace_bytes, ws.bytes));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned char* base = static_cast<const unsigned char*>(packed);
  // model.py:497-503,515  LSTMCell(embedding(previous_word)).  Without dropout on the embedding (GloVe: model.py:499-502)
  // the input half of the gates depends on the word id only: it is read from the per-token table built at pack time and
  // only W_hh h_0 is multiplied per step; otherwise the embedding rows are gathered + split on the fly
  {
    PkParams q{};
    const bool hoist = drop_e == nullptr && !g_disable_hoist;
    q.a_pk = base + (hoist ? P.a_hh : P.a_gates); q.nkb = hoist ? P.nkb_h : P.nkb_gates;
    if (hoist) {
      q.g.nseg = 1;
      q.g.seg[0] = GemmSeg{h0, H, nullptr, nullptr, 0, nullptr, 0, H, 0};
    } else {
      q.g.nseg = 2;
      q.g.seg[0] = GemmSeg{w->embedding, Ew, prev_word, drop_e, drop_e ? Ew : 0, nullptr, 0, Ew, 0};
      q.g.seg[1] = GemmSeg{h0, H, nullptr, nullptr, 0, nullptr, 0, H, 0};
    }
    q.g.M = B; q.g.N = 4 * H;
    LstmEpilogue& e = q.g.lstm;
    if (hoist) {
      e.addend = reinterpret_cast<const float*>(base + P.tdec); e.ld_addend = 4 * H; e.addend_rows = prev_word;
    }
    e.H = H; e.b_ih = w->lstm_b_ih; e.b_hh = w->lstm_b_hh; e.c0 = c0; e.drop_h = drop_h;
    e.h1 = h1; e.c1 = c1; e.h1_drop = ws.h1d; e.gates_act = ws.gates_act;
    SFB_PROPAGATE(launch_gemm_pk(q, st, ws.pk, ws.pk_bytes));
  }
  auto proj = [&](const unsigned char* a_pk, const float* x, int n_out, float* out, int ldo, const float* bias,
                  const float* padd, int ld_padd, int act) {
    PkParams q{};
    q.a_pk = a_pk; q.nkb = P.nkb_h;
    q.g.nseg = 1;
    q.g.seg[0] = GemmSeg{x, H, nullptr, nullptr, 0, nullptr, 0, H, 0};
    q.g.M = B; q.g.N = n_out; q.g.out = out; q.g.ldo = ldo; q.g.bias0 = bias; q.g.padd = padd; q.g.ld_padd = ld_padd; q.g.act = act;
    return launch_gemm_pk(q, st, ws.pk, ws.pk_bytes);
  };
  // model.py:516-517  SoftDotAttention(drop(h_1), ctx, path_mask)
  SFB_PROPAGATE(proj(base + P.a_th, ws.h1d, 2 * H, ws.th, 2 * H, nullptr, nullptr, 0, 0));
  {
    AttnParams a{};
    a.q = ws.th; a.ldq = 2 * H; a.R = T; a.D = H;
    a.segA = ctx; a.strideA_b = (long long)T * H; a.strideA_r = H; a.lenA = H; a.lenB = 0;
    a.mask = ctx_mask; a.ldmask = T;
    a.out = ws.wc; a.ldo = H; a.alpha = alpha; a.ldalpha = T;
    SFB_PROPAGATE(launch_soft_dot_attention(a, B, ws.at, ws.at_bytes, st));
  }
  SFB_PROPAGATE(proj(base + P.a_wc, ws.wc, H, ws.htilde, H, nullptr, ws.th + H, 2 * H, 1));
  // model.py:518  logit = decoder2action(h_tilde)
  return proj(base + P.a_voc, ws.htilde, vocab, logit, vocab, w->b_voc, nullptr, 0, 0);
}

}  // extern "C"

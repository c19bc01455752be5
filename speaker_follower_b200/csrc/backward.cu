// backward.cu — hand-written backward kernels of one follower decode step (autograd of AttnDecoderLSTM.forward,
// model.py:377-397, as driven by follower.py:1001-1020).  The dense products reuse the skinny GEMM (gemm_simt.cu);
// this file holds what is not a plain GEMM:
//   score_bwd_kernel      d(logit) -> dg[b,:] = sum_a dlogit[b,a] u_{b,a},  dsum[b] = sum_a dlogit[b,a]   (EltwiseProdScoring)
//   text_attn_bwd_kernel  softmax / weighted-sum backward of SoftDotAttention over ctx, one CTA per batch element:
//                         ctx rows streamed once into shared memory, d(ctx) written, dt and the weighted context emitted
//   vis_attn_bwd_kernel   the same for VisualSoftDotAttention over the 36-view slab (read once from the feature table):
//                         dq = sum_i alpha_i (dfeat . V_i - dfeat . f) V_i
//   lstm_cell_bwd_kernel  nn.LSTMCell pointwise backward from the saved activated gates
//   outer_accum_kernel    weight gradients dW[n,k] (+)= sum_b Y[b,n] X[b,k] (reduction over the batch), exact fp32
//   colsum_kernel         bias gradients
#include "kernels.h"

namespace sfb {

// ---------------------------------------------------------------- scoring backward
// one CTA per batch element; candidates dense [B,A,E] or gathered from the feature table (env.py:60-75)
__global__ void __launch_bounds__(256) score_bwd_kernel(const ScoreBwdParams p) {
  const int b = blockIdx.x, tid = threadIdx.x;
  __shared__ float dl[64];
  for (int a = tid; a < p.A; a += 256) dl[a] = p.dlogit[(size_t)b * p.A + a];
  __syncthreads();
  float s = 0.f;
  for (int a = 0; a < p.A; ++a) s += dl[a];
  if (tid == 0) p.dsum[b] = s;
  const int nvec = p.E >> 2;
  const int loc = p.cand_table ? p.E - p.img_dim : 0, grp = loc >> 2;
  for (int j = tid; j < nvec; j += 256) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < p.A; ++a) {
      const float w = dl[a];
      if (w == 0.f) continue;
      float4 u;
      if (p.cand_table == nullptr) {
        u = *reinterpret_cast<const float4*>(p.all_u_t + ((size_t)b * p.A + a) * p.E + 4 * j);
      } else {
        const int v = p.cand_view[(size_t)b * p.A + a];
        if (v < 0) continue;
        const int k = 4 * j;
        if (k < p.img_dim) {
          u = *reinterpret_cast<const float4*>(p.cand_table + ((size_t)p.vp_idx[b] * p.cand_V + v) * p.img_dim + k);
        } else {
          const float t = p.cand_trig[((size_t)b * p.A + a) * 4 + (k - p.img_dim) / grp];
          u = make_float4(t, t, t, t);
        }
      }
      acc.x = fmaf(w, u.x, acc.x); acc.y = fmaf(w, u.y, acc.y); acc.z = fmaf(w, u.z, acc.z); acc.w = fmaf(w, u.w, acc.w);
    }
    *reinterpret_cast<float4*>(p.dg + (size_t)b * p.E + 4 * j) = acc;
  }
}

int32_t launch_score_bwd(const ScoreBwdParams& p, cudaStream_t st) {
  SFB_CHECK_ARG(p.A <= 64 && (p.E % 4) == 0, "score_bwd: A <= 64, E % 4 == 0");
  SFB_CHECK_ARG(!p.cand_table || (((p.E - p.img_dim) % 16) == 0 && (p.img_dim % 4) == 0), "score_bwd: bad gather source");
  score_bwd_kernel<<<p.B, 256, 0, st>>>(p);
  SFB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------- pointwise pieces
// dz = dh~ (1 - h~^2)   (tanh of SoftDotAttention.linear_out, model.py:142)
__global__ void tanh_bwd_kernel(const float* dy, const float* y, float* dz, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float t = y[i]; dz[i] = dy[i] * (1.f - t * t); }
}
// r = w_o (.) th ; used as the left factor of dW_a / db_a          (model.py:348-351)
// dth = w_o (.) (v + b_a dsum) ; dw_o partial products th (.) (v + b_a dsum)
__global__ void scoring_mid_kernel(const float* th, const float* v, const float* w_o, const float* b_a, const float* dsum,
                                   float* r, float* dth, float* prod, int B, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  const float vv = v[i] + b_a[d] * dsum[b];
  r[i] = w_o[d] * th[i];
  dth[i] = w_o[d] * vv;
  prod[i] = th[i] * vv;
}
int32_t launch_tanh_bwd(const float* dy, const float* y, float* dz, int n, cudaStream_t st) {
  tanh_bwd_kernel<<<(n + 255) / 256, 256, 0, st>>>(dy, y, dz, n);
  SFB_CHECK_LAUNCH();
  return 0;
}
int32_t launch_scoring_mid(const float* th, const float* v, const float* w_o, const float* b_a, const float* dsum, float* r,
                           float* dth, float* prod, int B, int D, cudaStream_t st) {
  scoring_mid_kernel<<<(B * D + 255) / 256, 256, 0, st>>>(th, v, w_o, b_a, dsum, r, dth, prod, B, D);
  SFB_CHECK_LAUNCH();
  return 0;
}

// nn.LSTMCell backward (model.py:393) from the activated gates saved by the forward epilogue
//   dh = g_h1 + dh1d (.) drop_h ;  do = dh tanh(c1) ; dc = g_c1 + dh o (1 - tanh(c1)^2)
//   di = dc g ; df = dc c0 ; dg = dc i ; dc0 = dc f ; pre-activation grads: i(1-i), f(1-f), 1-g^2, o(1-o)
__global__ void lstm_cell_bwd_kernel(const LstmBwdParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.B * p.H) return;
  const int b = idx / p.H, j = idx - b * p.H;
  const float* ga = p.gates_act + (size_t)b * 4 * p.H + j;
  const float ig = ga[0], fg = ga[p.H], gt = ga[2 * p.H], og = ga[3 * p.H];
  float dh = p.g_h1 ? p.g_h1[idx] : 0.f;
  if (p.dh1d) dh += p.dh1d[idx] * (p.drop_h ? p.drop_h[idx] : 1.f);
  const float tc = tanhf(p.c1[idx]);
  const float dc = (p.g_c1 ? p.g_c1[idx] : 0.f) + dh * og * (1.f - tc * tc);
  float* dg = p.dgates + (size_t)b * 4 * p.H + j;
  dg[0] = dc * gt * ig * (1.f - ig);
  dg[p.H] = dc * p.c0[idx] * fg * (1.f - fg);
  dg[2 * p.H] = dc * ig * (1.f - gt * gt);
  dg[3 * p.H] = dh * tc * og * (1.f - og);
  p.dc0[idx] = dc * fg;
}
int32_t launch_lstm_cell_bwd(const LstmBwdParams& p, cudaStream_t st) {
  lstm_cell_bwd_kernel<<<(p.B * p.H + 255) / 256, 256, 0, st>>>(p);
  SFB_CHECK_LAUNCH();
  return 0;
}

// EncoderLSTM BPTT (model.py:81-104 under autograd): one cell step of the packed sequence
__global__ void lstm_seq_bwd_kernel(const LstmSeqBwdParams p) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.B * p.H) return;
  const int b = idx / p.H, j = idx - b * p.H;
  float* dg = p.dgates + (size_t)b * 4 * p.H + j;
  if (p.t >= p.lengths[b]) {      // the row ended before this step: its state was carried, nothing was emitted
    dg[0] = 0.f; dg[p.H] = 0.f; dg[2 * p.H] = 0.f; dg[3 * p.H] = 0.f;
    p.dc_prev[idx] = p.dc_in[idx];
    p.dh_pass[idx] = p.dh_in[idx];
    return;
  }
  const float* ga = p.gates_act + (size_t)b * 4 * p.H + j;
  const float ig = ga[0], fg = ga[p.H], gt = ga[2 * p.H], og = ga[3 * p.H];
  const float dh = p.dh_in[idx] + (p.g_out ? p.g_out[(size_t)b * p.ld_g_out + j] : 0.f);
  const float tc = tanhf(p.c_cur[idx]);
  const float dc = p.dc_in[idx] + dh * og * (1.f - tc * tc);
  dg[0] = dc * gt * ig * (1.f - ig);
  dg[p.H] = dc * p.c_prev[idx] * fg * (1.f - fg);
  dg[2 * p.H] = dc * ig * (1.f - gt * gt);
  dg[3 * p.H] = dh * tc * og * (1.f - og);
  p.dc_prev[idx] = dc * fg;
  p.dh_pass[idx] = 0.f;
}
int32_t launch_lstm_seq_bwd(const LstmSeqBwdParams& p, cudaStream_t st) {
  lstm_seq_bwd_kernel<<<(p.B * p.H + 255) / 256, 256, 0, st>>>(p);
  SFB_CHECK_LAUNCH();
  return 0;
}
// out[(t*B + b), :] = embedding[seq[b, t], :] (.) drop[(b*maxlen + t), :]   (the x_t rows, in the tape's [t][b] order)
__global__ void gather_embed_kernel(const float* emb, int Ew, const int32_t* seq, const float* drop, float* out, int B, int maxlen) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * maxlen * Ew) return;
  const int k = (int)(i % Ew);
  const long long r = i / Ew;
  const int b = (int)(r % B), t = (int)(r / B);
  const size_t m = (size_t)b * maxlen + t;
  float v = emb[(size_t)seq[m] * Ew + k];
  if (drop) v *= drop[m * Ew + k];
  out[i] = v;
}
int32_t launch_gather_embed(const float* emb, int Ew, const int32_t* seq, const float* drop, float* out, int B, int maxlen, cudaStream_t st) {
  const long long n = (long long)B * maxlen * Ew;
  gather_embed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(emb, Ew, seq, drop, out, B, maxlen);
  SFB_CHECK_LAUNCH();
  return 0;
}

// x = [u_prev | feat] (.) drop_x   (the LSTM input of model.py:391-392, needed as the right factor of dW_ih)
__global__ void assemble_x_kernel(const float* u, const float* f, const float* drop, float* x, int B, int E, int F) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, W = E + F;
  if (i >= B * W) return;
  const int b = i / W, k = i - b * W;
  const float v = k < E ? u[(size_t)b * E + k] : f[(size_t)b * F + k - E];
  x[i] = drop ? v * drop[i] : v;
}
int32_t launch_assemble_x(const float* u, const float* f, const float* drop, float* x, int B, int E, int F, cudaStream_t st) {
  assemble_x_kernel<<<(B * (E + F) + 255) / 256, 256, 0, st>>>(u, f, drop, x, B, E, F);
  SFB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------- attention backward (one CTA per batch element)
// rows r_l (l < R, un-masked), weights alpha_l (forward softmax), upstream gradient dout of out = sum_l alpha_l r_l and
// the query qv of the scores s_l = r_l . qv:
//   dalpha_l = dout . r_l ;  c = sum_l alpha_l dalpha_l ;  ds_l = alpha_l (dalpha_l - c)
//   dq = sum_l ds_l r_l ;  (optional) drows_l = alpha_l dout + ds_l qv ;  (optional) wsum = sum_l alpha_l r_l
// The rows are fetched once (bulk async copies into shared memory when they fit, else re-read through L2).
template <int NT>
__global__ void __launch_bounds__(NT) attn_bwd_kernel(const AttnBwdParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* rows = reinterpret_cast<float*>(smem_raw);                      // [RS][D] staged rows (RS = staged count)
  float* dal = rows + (size_t)p.stage_rows * p.D;                       // [R] dalpha, then ds
  float* redbuf = dal + ((p.R + 3) & ~3);                               // [NT/32]
  uint64_t* bar = reinterpret_cast<uint64_t*>(redbuf + 32);
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NW = NT / 32;
  const int D = p.D, nvec = D >> 2;
  const uint8_t* mrow = p.mask ? p.mask + (size_t)b * p.ldmask : nullptr;
  const long long ia = p.idxA ? (long long)p.idxA[b] : (long long)b;
  const long long ib = p.idxB ? (long long)p.idxB[b] : (long long)b;
  const float* baseA = p.segA + (size_t)ia * p.strideA_b;
  const float* baseB = p.lenB > 0 ? p.segB + (size_t)ib * p.strideB_b : nullptr;
  const bool staged = p.stage_rows >= p.R;
  if (staged && tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    uint32_t bytes = 0;
    for (int l = 0; l < p.R; ++l)
      if (!(mrow && mrow[l])) bytes += (uint32_t)D * 4u;
    mbar_expect_tx(bar, bytes);
    for (int l = 0; l < p.R; ++l) {
      if (mrow && mrow[l]) continue;
      bulk_g2s(rows + (size_t)l * D, baseA + (size_t)l * p.strideA_r, (uint32_t)p.lenA * 4u, bar);
      if (p.lenB > 0) bulk_g2s(rows + (size_t)l * D + p.lenA, baseB + (size_t)l * p.strideB_r, (uint32_t)p.lenB * 4u, bar);
    }
  }
  __syncthreads();
  if (staged) mbar_wait(bar, 0);
  auto row_ptr = [&](int l, int k) -> const float* {   // address of element k of row l (staged or global)
    if (staged) return rows + (size_t)l * D + k;
    return k < p.lenA ? baseA + (size_t)l * p.strideA_r + k : baseB + (size_t)l * p.strideB_r + (k - p.lenA);
  };
  const float* dout = p.dout + (size_t)b * p.lddout;
  const float* alpha = p.alpha + (size_t)b * p.ldalpha;
  // pass 1: dalpha_l = dout . r_l  (one warp per row)
  for (int l = warp; l < p.R; l += NW) {
    float acc = 0.f;
    if (!(mrow && mrow[l]))
      for (int j = lane; j < nvec; j += 32) {
        const float4 r = *reinterpret_cast<const float4*>(row_ptr(l, 4 * j));
        float4 d = *reinterpret_cast<const float4*>(dout + 4 * j);
        if (p.dout_scale) {
          const float4 s = *reinterpret_cast<const float4*>(p.dout_scale + (size_t)b * p.ldscale + 4 * j);
          d.x *= s.x; d.y *= s.y; d.z *= s.z; d.w *= s.w;
        }
        acc = fmaf(r.x, d.x, acc); acc = fmaf(r.y, d.y, acc); acc = fmaf(r.z, d.z, acc); acc = fmaf(r.w, d.w, acc);
      }
    acc = warp_sum(acc);
    if (lane == 0) dal[l] = acc;
  }
  __syncthreads();
  // c = sum alpha dalpha ; ds_l
  float part = 0.f;
  for (int l = tid; l < p.R; l += NT) part += (mrow && mrow[l]) ? 0.f : alpha[l] * dal[l];
  part = warp_sum(part);
  if (lane == 0) redbuf[warp] = part;
  __syncthreads();
  float c = 0.f;
  for (int w = 0; w < NW; ++w) c += redbuf[w];
  __syncthreads();
  for (int l = tid; l < p.R; l += NT) dal[l] = (mrow && mrow[l]) ? 0.f : alpha[l] * (dal[l] - c);
  __syncthreads();
  // pass 2: column-wise — dq = sum ds_l r_l, wsum = sum alpha_l r_l, drows_l = alpha_l dout + ds_l qv
  for (int j = tid; j < nvec; j += NT) {
    float4 dq = make_float4(0.f, 0.f, 0.f, 0.f), ws = dq;
    float4 d = *reinterpret_cast<const float4*>(dout + 4 * j);
    if (p.dout_scale) {
      const float4 s = *reinterpret_cast<const float4*>(p.dout_scale + (size_t)b * p.ldscale + 4 * j);
      d.x *= s.x; d.y *= s.y; d.z *= s.z; d.w *= s.w;
    }
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.drows) qv = *reinterpret_cast<const float4*>(p.qv + (size_t)b * p.ldq + 4 * j);
    for (int l = 0; l < p.R; ++l) {
      const bool m = mrow && mrow[l];
      const float a = m ? 0.f : alpha[l], ds = dal[l];
      if (!m) {
        const float4 r = *reinterpret_cast<const float4*>(row_ptr(l, 4 * j));
        dq.x = fmaf(ds, r.x, dq.x); dq.y = fmaf(ds, r.y, dq.y); dq.z = fmaf(ds, r.z, dq.z); dq.w = fmaf(ds, r.w, dq.w);
        ws.x = fmaf(a, r.x, ws.x); ws.y = fmaf(a, r.y, ws.y); ws.z = fmaf(a, r.z, ws.z); ws.w = fmaf(a, r.w, ws.w);
      }
      if (p.drows) {
        float4 o;
        o.x = fmaf(a, d.x, ds * qv.x); o.y = fmaf(a, d.y, ds * qv.y); o.z = fmaf(a, d.z, ds * qv.z); o.w = fmaf(a, d.w, ds * qv.w);
        *reinterpret_cast<float4*>(p.drows + ((size_t)b * p.R + l) * D + 4 * j) = o;
      }
    }
    *reinterpret_cast<float4*>(p.dq + (size_t)b * p.lddq + 4 * j) = dq;
    if (p.wsum) *reinterpret_cast<float4*>(p.wsum + (size_t)b * p.ldwsum + 4 * j) = ws;
  }
}

int32_t launch_attn_bwd(AttnBwdParams p, int B, cudaStream_t st) {
  SFB_CHECK_ARG((p.D % 4) == 0 && p.lenA + p.lenB == p.D && (p.lenA % 4) == 0 && (p.lenB % 4) == 0, "attn_bwd: bad row layout");
  SFB_CHECK_ARG(p.R >= 1 && p.dout && p.alpha && p.dq, "attn_bwd: NULL argument");
  const size_t row_bytes = (size_t)p.D * 4;
  const size_t fixed = (((size_t)p.R + 3) & ~size_t(3)) * 4 + 32 * 4 + 16;
  p.stage_rows = ((size_t)p.R * row_bytes + fixed <= 200 * 1024) ? p.R : 0;   // slabs (313 KB) are re-read through L2 in pass 2
  const size_t smem = (size_t)p.stage_rows * row_bytes + fixed;
  static SmemMarks marks;
  SFB_CHECK_CUDA(ensure_dynamic_smem(attn_bwd_kernel<256>, smem, marks));
  attn_bwd_kernel<256><<<B, 256, smem, st>>>(p);
  SFB_CHECK_LAUNCH();
  return 0;
}

// ---------------------------------------------------------------- weight gradients: out[n,k] (+)= sum_b Y[b,n] X[b,k]
// 64 x 64 output tile per CTA, the batch reduced in chunks of 32 rows staged in shared memory; exact fp32.
__global__ void __launch_bounds__(256) outer_accum_kernel(const OuterParams p) {
  __shared__ float Ys[32][64 + 4];
  __shared__ float Xs[32][64 + 4];
  const int n0 = blockIdx.y * 64, k0 = blockIdx.x * 64, tid = threadIdx.x;
  const int tn = (tid >> 4) * 4, tk = (tid & 15) * 4;     // each thread: 4 x 4 outputs
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int b0 = 0; b0 < p.B; b0 += 32) {
    for (int i = tid; i < 32 * 64; i += 256) {
      const int bb = i >> 6, c = i & 63, b = b0 + bb;
      Ys[bb][c] = (b < p.B && n0 + c < p.N) ? p.Y[(size_t)b * p.ldy + n0 + c] : 0.f;
      float xv = 0.f;
      if (b < p.B && k0 + c < p.K) {
        xv = p.X[(size_t)b * p.ldx + k0 + c];
      }
      Xs[bb][c] = xv;
    }
    __syncthreads();
#pragma unroll 8
    for (int bb = 0; bb < 32; ++bb) {
      const float4 y = *reinterpret_cast<const float4*>(&Ys[bb][tn]);
      const float4 x = *reinterpret_cast<const float4*>(&Xs[bb][tk]);
      const float yy[4] = {y.x, y.y, y.z, y.w}, xx[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(yy[i], xx[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + tn + i;
    if (n >= p.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tk + j;
      if (k >= p.K) continue;
      float* o = p.out + (size_t)n * p.ldo + k;
      *o = p.accumulate ? *o + acc[i][j] : acc[i][j];
    }
  }
}
int32_t launch_outer_accum(const OuterParams& p, cudaStream_t st) {
  SFB_CHECK_ARG(p.Y && p.X && p.out && p.B >= 1 && p.N >= 1 && p.K >= 1, "outer_accum: bad arguments");
  outer_accum_kernel<<<dim3((p.K + 63) / 64, (p.N + 63) / 64), 256, 0, st>>>(p);
  SFB_CHECK_LAUNCH();
  return 0;
}

// out[n] (+)= sum_b Y[b,n] (* optional scale)
__global__ void colsum_kernel(const float* Y, int ldy, int B, int N, float* out, int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += Y[(size_t)b * ldy + n];
  out[n] = accumulate ? out[n] + s : s;
}
int32_t launch_colsum(const float* Y, int ldy, int B, int N, float* out, int accumulate, cudaStream_t st) {
  colsum_kernel<<<(N + 127) / 128, 128, 0, st>>>(Y, ldy, B, N, out, accumulate);
  SFB_CHECK_LAUNCH();
  return 0;
}

}  // namespace sfb

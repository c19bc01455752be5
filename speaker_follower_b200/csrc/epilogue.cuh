// epilogue.cuh — GEMM epilogues shared by the FFMA (gemm_simt.cu) and tensor-core (gemm_tc.cu) kernels.
#pragma once
#include "kernels.h"

namespace sfb {

// LSTM cell update for one (row, hidden unit); `g` are the four pre-activation gate sums WITHOUT biases.
__device__ __forceinline__ void lstm_update(const GemmParams& p, int m, int unit, float gi, float gf, float gg,
                                            float go) {
  const LstmEpilogue& e = p.lstm;
  const int H = e.H;
  const size_t idx = (size_t)m * H + unit;
  if (e.lengths && e.t >= e.lengths[m]) {   // packed sequence: this row has ended, carry the state
    e.c1[idx] = e.c0[idx];
    e.h1[idx] = e.h0[idx];
    if (e.seq_out) e.seq_out[(size_t)m * e.ld_seq_out + unit] = 0.f;
    return;
  }
  gi += __ldg(e.b_ih + unit) + __ldg(e.b_hh + unit);
  gf += __ldg(e.b_ih + H + unit) + __ldg(e.b_hh + H + unit);
  gg += __ldg(e.b_ih + 2 * H + unit) + __ldg(e.b_hh + 2 * H + unit);
  go += __ldg(e.b_ih + 3 * H + unit) + __ldg(e.b_hh + 3 * H + unit);
  if (e.addend) {
    const float* a = e.addend + (size_t)m * e.ld_addend + unit;
    gi += a[0]; gf += a[H]; gg += a[2 * H]; go += a[3 * H];
  }
  const float ig = sigmoidf_acc(gi), fg = sigmoidf_acc(gf), gt = tanhf(gg), og = sigmoidf_acc(go);
  const float c1 = fg * e.c0[idx] + ig * gt;
  const float h1 = og * tanhf(c1);
  e.c1[idx] = c1;
  e.h1[idx] = h1;
  if (e.h1_drop) e.h1_drop[idx] = e.drop_h ? h1 * e.drop_h[idx] : h1;
  if (e.seq_out) e.seq_out[(size_t)m * e.ld_seq_out + unit] = h1;
  if (e.gates_act) {
    float* ga = e.gates_act + (size_t)m * 4 * H + unit;
    ga[0] = ig; ga[H] = fg; ga[2 * H] = gt; ga[3 * H] = og;
  }
}

// four consecutive hidden units (unit % 4 == 0): 128-bit loads/stores when H % 4 == 0 and no sequence mode
__device__ __forceinline__ void lstm_update4(const GemmParams& p, int m, int unit, float4 gi, float4 gf, float4 gg, float4 go) {
  const LstmEpilogue& e = p.lstm;
  const int H = e.H;
  if (e.lengths || e.addend || e.seq_out || (H & 3)) {
    lstm_update(p, m, unit + 0, gi.x, gf.x, gg.x, go.x);
    lstm_update(p, m, unit + 1, gi.y, gf.y, gg.y, go.y);
    lstm_update(p, m, unit + 2, gi.z, gf.z, gg.z, go.z);
    lstm_update(p, m, unit + 3, gi.w, gf.w, gg.w, go.w);
    return;
  }
  auto ld4 = [](const float* q) { return __ldg(reinterpret_cast<const float4*>(q)); };
  const float4 bi0 = ld4(e.b_ih + unit), bi1 = ld4(e.b_hh + unit), bf0 = ld4(e.b_ih + H + unit), bf1 = ld4(e.b_hh + H + unit);
  const float4 bg0 = ld4(e.b_ih + 2 * H + unit), bg1 = ld4(e.b_hh + 2 * H + unit), bo0 = ld4(e.b_ih + 3 * H + unit),
               bo1 = ld4(e.b_hh + 3 * H + unit);
  const size_t idx = (size_t)m * H + unit;
  const float4 c0 = *reinterpret_cast<const float4*>(e.c0 + idx);
  float4 dh = make_float4(1.f, 1.f, 1.f, 1.f);
  if (e.h1_drop && e.drop_h) dh = *reinterpret_cast<const float4*>(e.drop_h + idx);
  const float pi[4] = {gi.x + bi0.x + bi1.x, gi.y + bi0.y + bi1.y, gi.z + bi0.z + bi1.z, gi.w + bi0.w + bi1.w};
  const float pf[4] = {gf.x + bf0.x + bf1.x, gf.y + bf0.y + bf1.y, gf.z + bf0.z + bf1.z, gf.w + bf0.w + bf1.w};
  const float pg[4] = {gg.x + bg0.x + bg1.x, gg.y + bg0.y + bg1.y, gg.z + bg0.z + bg1.z, gg.w + bg0.w + bg1.w};
  const float po[4] = {go.x + bo0.x + bo1.x, go.y + bo0.y + bo1.y, go.z + bo0.z + bo1.z, go.w + bo0.w + bo1.w};
  const float cc[4] = {c0.x, c0.y, c0.z, c0.w};
  float ig[4], fg[4], gt[4], og[4], c1[4], h1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ig[i] = sigmoidf_acc(pi[i]); fg[i] = sigmoidf_acc(pf[i]); gt[i] = tanhf(pg[i]); og[i] = sigmoidf_acc(po[i]);
    c1[i] = fg[i] * cc[i] + ig[i] * gt[i];
    h1[i] = og[i] * tanhf(c1[i]);
  }
  *reinterpret_cast<float4*>(e.c1 + idx) = make_float4(c1[0], c1[1], c1[2], c1[3]);
  *reinterpret_cast<float4*>(e.h1 + idx) = make_float4(h1[0], h1[1], h1[2], h1[3]);
  if (e.h1_drop) *reinterpret_cast<float4*>(e.h1_drop + idx) = make_float4(h1[0] * dh.x, h1[1] * dh.y, h1[2] * dh.z, h1[3] * dh.w);
  if (e.gates_act) {
    float* ga = e.gates_act + (size_t)m * 4 * H + unit;
    *reinterpret_cast<float4*>(ga) = make_float4(ig[0], ig[1], ig[2], ig[3]);
    *reinterpret_cast<float4*>(ga + H) = make_float4(fg[0], fg[1], fg[2], fg[3]);
    *reinterpret_cast<float4*>(ga + 2 * H) = make_float4(gt[0], gt[1], gt[2], gt[3]);
    *reinterpret_cast<float4*>(ga + 3 * H) = make_float4(og[0], og[1], og[2], og[3]);
  }
}

__device__ __forceinline__ void plain_store(const GemmParams& p, int m, int n, float v) {
  if (p.out2 && n >= p.n_split) {
    const int n2 = n - p.n_split;
    if (p.bias2) v += __ldg(p.bias2 + n2);
    p.out2[(size_t)m * p.ldo2 + n2] = v;
    return;
  }
  if (p.out2 && p.n1_valid > 0 && n >= p.n1_valid) return;   // padding rows of the first output block
  if (p.bias0) v += __ldg(p.bias0 + n);
  if (p.bias1) v += __ldg(p.bias1 + n);
  if (p.padd) v += p.padd[(size_t)m * p.ld_padd + n];
  if (p.act == 1) v = tanhf(v);
  if (p.oscale) v *= __ldg(p.oscale + n);
  p.out[(size_t)m * p.ldo + n] = v;
}

// four consecutive output features n..n+3 of row m (n % 4 == 0); vector store when the row stride allows it
__device__ __forceinline__ void plain_store4(const GemmParams& p, int m, int n, float4 v) {
  if (p.out2 && n >= p.n_split) {   // second output (n_split is tile aligned, so a float4 never straddles)
    const int n2 = n - p.n_split;
    if (n + 3 < p.N && (p.ldo2 & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out2) & 15u) == 0) {
      if (p.bias2) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias2 + n2));
        v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w;
      }
      *reinterpret_cast<float4*>(p.out2 + (size_t)m * p.ldo2 + n2) = v;
    } else {
      if (n < p.N) plain_store(p, m, n, v.x);
      if (n + 1 < p.N) plain_store(p, m, n + 1, v.y);
      if (n + 2 < p.N) plain_store(p, m, n + 2, v.z);
      if (n + 3 < p.N) plain_store(p, m, n + 3, v.w);
    }
    return;
  }
  const int n_end = (p.out2 && p.n1_valid > 0) ? p.n1_valid : p.N;
  if (n + 3 < n_end && (p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15u) == 0) {
    float r[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float x = r[i];
      if (p.bias0) x += __ldg(p.bias0 + n + i);
      if (p.bias1) x += __ldg(p.bias1 + n + i);
      if (p.padd) x += p.padd[(size_t)m * p.ld_padd + n + i];
      if (p.act == 1) x = tanhf(x);
      if (p.oscale) x *= __ldg(p.oscale + n + i);
      r[i] = x;
    }
    *reinterpret_cast<float4*>(p.out + (size_t)m * p.ldo + n) = make_float4(r[0], r[1], r[2], r[3]);
  } else {
    if (n < p.N) plain_store(p, m, n, v.x);
    if (n + 1 < p.N) plain_store(p, m, n + 1, v.y);
    if (n + 2 < p.N) plain_store(p, m, n + 2, v.z);
    if (n + 3 < p.N) plain_store(p, m, n + 3, v.w);
  }
}

}  // namespace sfb

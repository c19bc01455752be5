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

__device__ __forceinline__ void plain_store(const GemmParams& p, int m, int n, float v) {
  if (p.bias0) v += __ldg(p.bias0 + n);
  if (p.bias1) v += __ldg(p.bias1 + n);
  if (p.padd) v += p.padd[(size_t)m * p.ld_padd + n];
  if (p.act == 1) v = tanhf(v);
  if (p.oscale) v *= __ldg(p.oscale + n);
  p.out[(size_t)m * p.ldo + n] = v;
}

// four consecutive output features n..n+3 of row m (n % 4 == 0); vector store when the row stride allows it
__device__ __forceinline__ void plain_store4(const GemmParams& p, int m, int n, float4 v) {
  if (n + 3 < p.N && (p.ldo & 3) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15u) == 0) {
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

// epilogue.cuh — GEMM epilogues shared by the FFMA (gemm_simt.cu) and tensor-core (gemm_tc.cu) kernels.
#pragma once
#include <cuda_bf16.h>

#include "kernels.h"

namespace sfb {

// h1 as bf16 (hi, lo) into the packed activation operand of the next step's gate GEMM (layout: pack.cu)
__device__ __forceinline__ void lstm_pack_to(unsigned char* base, int nkb, int kb0, int NB, int rows_per_z, int m, int unit, float h1) {
  const int zt = m / rows_per_z, r = m - zt * rows_per_z;
  const size_t half = (size_t)NB * 128;
  const __nv_bfloat16 hi = __float2bfloat16_rn(h1);
  const __nv_bfloat16 lo = __float2bfloat16_rn(h1 - __bfloat162float(hi));
  unsigned char* dst = base + ((size_t)zt * nkb + kb0 + (unit >> 6)) * (2 * half) + (size_t)(r >> 3) * 1024 +
                       (size_t)((unit & 63) >> 3) * 128 + (size_t)(r & 7) * 16 + (size_t)(unit & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(dst) = hi;
  *reinterpret_cast<__nv_bfloat16*>(dst + half) = lo;
}
__device__ __forceinline__ void lstm_pack_h(const LstmEpilogue& e, int m, int unit, float h1, float h1d) {
  if (e.hpk) lstm_pack_to(e.hpk, e.hpk_nkb, e.hpk_kb0, e.hpk_NB, e.hpk_rows_per_z, m, unit, h1);
  if (e.hdpk) lstm_pack_to(e.hdpk, e.H / 64, 0, e.hpk_NB, e.hpk_rows_per_z, m, unit, h1d);
}

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
    const float* a = e.addend + (size_t)(e.addend_rows ? e.addend_rows[m] : m) * e.ld_addend + unit;
    gi += a[0]; gf += a[H]; gg += a[2 * H]; go += a[3 * H];
  }
  const float ig = sigmoidf_acc(gi), fg = sigmoidf_acc(gf), gt = tanhf(gg), og = sigmoidf_acc(go);
  const float c1 = fg * e.c0[idx] + ig * gt;
  const float h1 = og * tanhf(c1);
  e.c1[idx] = c1;
  e.h1[idx] = h1;
  if (e.h1_drop) e.h1_drop[idx] = e.drop_h ? h1 * e.drop_h[idx] : h1;
  if (e.hpk || e.hdpk) lstm_pack_h(e, m, unit, h1, e.drop_h ? h1 * e.drop_h[idx] : h1);
  if (e.seq_out) e.seq_out[(size_t)m * e.ld_seq_out + unit] = h1;
  if (e.gates_act) {
    float* ga = e.gates_act + (size_t)m * 4 * H + unit;
    ga[0] = ig; ga[H] = fg; ga[2 * H] = gt; ga[3 * H] = og;
  }
}

__device__ __forceinline__ void plain_store(const GemmParams& p, int m, int n, float v) {
  const int mo = p.out_rows ? __ldg(p.out_rows + m) : m;
  if (p.out2 && n >= p.n_split) {
    const int n2 = n - p.n_split;
    if (p.bias2) v += __ldg(p.bias2 + n2);
    p.out2[(size_t)mo * p.ldo2 + n2] = v;
    return;
  }
  if (p.out2 && p.n1_valid > 0 && n >= p.n1_valid) return;   // padding rows of the first output block
  if (p.bias0) v += __ldg(p.bias0 + n);
  if (p.bias1) v += __ldg(p.bias1 + n);
  if (p.padd) v += p.padd[(size_t)m * p.ld_padd + n];
  if (p.act == 1) v = tanhf(v);
  if (p.oscale) v *= __ldg(p.oscale + n);
  p.out[(size_t)mo * p.ldo + n] = v;
}

// Epilogue operands of one float4 group, fetched BEFORE the split-K barrier so that their latency hides behind it.
struct PlainPre { float4 add; bool ok; };
__device__ __forceinline__ PlainPre plain_preload(const GemmParams& p, int m, int n) {
  PlainPre r;
  r.add = make_float4(0.f, 0.f, 0.f, 0.f);
  r.ok = false;
  const bool second = p.out2 && n >= p.n_split;
  const int n_end = second ? p.N : ((p.out2 && p.n1_valid > 0) ? p.n1_valid : p.N);
  if (n + 3 >= n_end || p.bias1 || p.oscale) return r;   // rare shapes take the plain path
  auto ld4 = [](const float* q) { return *reinterpret_cast<const float4*>(q); };
  if (second) {
    if (p.bias2) {
      if (reinterpret_cast<uintptr_t>(p.bias2) & 15u) return r;
      r.add = ld4(p.bias2 + (n - p.n_split));
    }
  } else {
    if (p.bias0) {
      if (reinterpret_cast<uintptr_t>(p.bias0) & 15u) return r;
      r.add = ld4(p.bias0 + n);
    }
    if (p.padd) {
      if ((reinterpret_cast<uintptr_t>(p.padd) & 15u) || (p.ld_padd & 3)) return r;
      const float4 a = ld4(p.padd + (size_t)m * p.ld_padd + n);
      r.add.x += a.x; r.add.y += a.y; r.add.z += a.z; r.add.w += a.w;
    }
  }
  r.ok = true;
  return r;
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
      *reinterpret_cast<float4*>(p.out2 + (size_t)(p.out_rows ? __ldg(p.out_rows + m) : m) * p.ldo2 + n2) = v;
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
    *reinterpret_cast<float4*>(p.out + (size_t)(p.out_rows ? __ldg(p.out_rows + m) : m) * p.ldo + n) = make_float4(r[0], r[1], r[2], r[3]);
  } else {
    if (n < p.N) plain_store(p, m, n, v.x);
    if (n + 1 < p.N) plain_store(p, m, n + 1, v.y);
    if (n + 2 < p.N) plain_store(p, m, n + 2, v.z);
    if (n + 3 < p.N) plain_store(p, m, n + 3, v.w);
  }
}

__device__ __forceinline__ void plain_store4_pre(const GemmParams& p, int m, int n, float4 v, const PlainPre& pre) {
  if (!pre.ok) { plain_store4(p, m, n, v); return; }
  const bool second = p.out2 && n >= p.n_split;
  v.x += pre.add.x; v.y += pre.add.y; v.z += pre.add.z; v.w += pre.add.w;
  if (!second && p.act == 1) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
  const int mo = p.out_rows ? __ldg(p.out_rows + m) : m;
  float* dst = second ? p.out2 + (size_t)mo * p.ldo2 + (n - p.n_split) : p.out + (size_t)mo * p.ldo + n;
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
    *reinterpret_cast<float4*>(dst) = v;
  } else {
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
  }
}

// one hidden unit per thread (all threads busy; operands fetched before the split-K barrier)
struct LstmPre1 { float bi, bf, bg, bo, c0, dh; bool ok; };
__device__ __forceinline__ LstmPre1 lstm_preload1(const GemmParams& p, int m, int unit) {
  const LstmEpilogue& e = p.lstm;
  const int H = e.H;
  LstmPre1 r;
  r.ok = !(e.lengths || e.addend || e.seq_out);
  if (!r.ok) return r;
  r.bi = __ldg(e.b_ih + unit) + __ldg(e.b_hh + unit);
  r.bf = __ldg(e.b_ih + H + unit) + __ldg(e.b_hh + H + unit);
  r.bg = __ldg(e.b_ih + 2 * H + unit) + __ldg(e.b_hh + 2 * H + unit);
  r.bo = __ldg(e.b_ih + 3 * H + unit) + __ldg(e.b_hh + 3 * H + unit);
  const size_t idx = (size_t)m * H + unit;
  r.c0 = e.c0[idx];
  r.dh = e.drop_h ? e.drop_h[idx] : 1.f;
  return r;
}
__device__ __forceinline__ void lstm_update1_pre(const GemmParams& p, int m, int unit, float gi, float gf, float gg, float go,
                                                 const LstmPre1& pre) {
  if (!pre.ok) { lstm_update(p, m, unit, gi, gf, gg, go); return; }
  const LstmEpilogue& e = p.lstm;
  const int H = e.H;
  const size_t idx = (size_t)m * H + unit;
  const float ig = sigmoidf_fast(gi + pre.bi), fg = sigmoidf_fast(gf + pre.bf), gt = tanhf_fast(gg + pre.bg), og = sigmoidf_fast(go + pre.bo);
  const float c1 = fg * pre.c0 + ig * gt;
  const float h1 = og * tanhf_fast(c1);
  e.c1[idx] = c1;
  e.h1[idx] = h1;
  if (e.h1_drop) e.h1_drop[idx] = h1 * pre.dh;
  if (e.hpk || e.hdpk) lstm_pack_h(e, m, unit, h1, h1 * pre.dh);
  if (e.gates_act) {
    float* ga = e.gates_act + (size_t)m * 4 * H + unit;
    ga[0] = ig; ga[H] = fg; ga[2 * H] = gt; ga[3 * H] = og;
  }
}

}  // namespace sfb

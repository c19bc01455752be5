// pack.cuh — one work item of the operand packer (see pack.cu): 8 consecutive K elements of one logical row ->
// one 16-byte bf16 "hi" and one 16-byte bf16 "lo" core-matrix row.
#pragma once
#include <cuda_bf16.h>

#include "kernels.h"

namespace sfb {

__device__ __forceinline__ void bf16_split8(const float4& a, const float4& b, uint4& hi, uint4& lo) {
  const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 hf = __bfloat1622float2(hh);
    const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// item idx in [0, ntile * nkb * R * 8): (tile, kblock, row, 8-wide k group).  Segments with x == NULL are skipped
// (their blocks are written by another producer, e.g. the attention kernel's epilogue).
__device__ __forceinline__ void pack_item(const PackParams& p, long long idx) {
  const size_t half = (size_t)p.R * 128;
  const int kc = (int)(idx & 7);
  long long t = idx >> 3;
  const int r = (int)(t % p.R);
  t /= p.R;
  const int kb = (int)(t % p.nkb);
  const int tile = (int)(t / p.nkb);
  int src_row;
  bool row_ok;
  if (p.lstm_H > 0) {   // gate-interleaved: tile row = gate*32 + unit_local
    src_row = (r >> 5) * p.lstm_H + tile * 32 + (r & 31);
    row_ok = tile * 32 + (r & 31) < p.lstm_H;
  } else {
    src_row = tile * p.rows_per_tile + r;
    row_ok = r < p.rows_per_tile && src_row < p.rows_valid;
  }
  int s = 0, cc = kb;
  while (s + 1 < p.nseg) {
    const int n = (p.seg[s].k + 63) / 64;
    if (cc < n) break;
    cc -= n;
    ++s;
  }
  const PackSeg& g = p.seg[s];
  if (g.x == nullptr) return;
  const int k = cc * 64 + kc * 8;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (row_ok && k < g.k) {
    const int xr = g.xrow ? g.xrow[src_row] : src_row;
    const float* src = g.x + (size_t)xr * g.ldx + k;
    a = *reinterpret_cast<const float4*>(src);
    if (k + 4 < g.k) b = *reinterpret_cast<const float4*>(src + 4);   // K % 8 == 4: the last group is half full
    if (g.xs) {
      const float* sp = g.xs + (size_t)src_row * g.ldxs + k;
      const float4 s0 = *reinterpret_cast<const float4*>(sp);
      const float4 s1 = (k + 4 < g.k) ? *reinterpret_cast<const float4*>(sp + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      a.x *= s0.x; a.y *= s0.y; a.z *= s0.z; a.w *= s0.w;
      b.x *= s1.x; b.y *= s1.y; b.z *= s1.z; b.w *= s1.w;
    }
  }
  uint4 hi, lo;
  bf16_split8(a, b, hi, lo);
  unsigned char* dst = p.out + ((size_t)tile * p.nkb + kb) * (2 * half) + (size_t)(r >> 3) * 1024 + (size_t)kc * 128 + (size_t)(r & 7) * 16;
  *reinterpret_cast<uint4*>(dst) = hi;
  *reinterpret_cast<uint4*>(dst + half) = lo;
}

}  // namespace sfb
